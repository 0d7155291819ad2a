"""2D equation plug-ins (module name = cfg['equation'], seistorch/model.py:64-66)."""
