"""Import alias: the package directory is `drmlt-mitsuba_b200/` (the hyphen is part of the
project name and not importable), so this module turns itself into that package."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "drmlt-mitsuba_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
