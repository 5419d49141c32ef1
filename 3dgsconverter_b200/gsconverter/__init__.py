"""Stand-alone stub of the ``gsconverter`` package root.

Only ``gsconverter.processing`` (the hot path) is re-implemented here; the reference's
orchestrator, CLI and format codecs are host code that is reused unchanged.  To drop this backend
into an installed 3dgsconverter, see INTEGRATION.md (copy ``processing/`` over the reference's, or
call ``gsx.dropin.patch()``).
"""
