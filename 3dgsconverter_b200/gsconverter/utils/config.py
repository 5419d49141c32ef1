"""Global debug switch (same name as the reference's utils/config.py:9 so ``main.py:302`` can set it)."""
DEBUG = False
