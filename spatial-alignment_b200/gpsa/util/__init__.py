from .util import *  # noqa: F401,F403  (mirrors reference gpsa/util/__init__.py:1)
