"""Console helpers with the reference's names (utils/utility_functions.py:15-29): messages go through
``tqdm.write`` so they do not tear the converter's progress bar."""
from . import config


def status_print(*args, **kwargs):
    text = " ".join(str(a) for a in args)
    try:
        from tqdm import tqdm
    except ImportError:
        print(text, **kwargs)
    else:
        tqdm.write(text, **kwargs)


def debug_print(*args, **kwargs):
    if config.DEBUG:
        status_print(*args, **kwargs)
