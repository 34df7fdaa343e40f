from .cif import *  # noqa: F401,F403  (same re-export as the reference's torch_cif/__init__.py)
