"""Import shim: makes the directory ``lwfbrook90.jl_b200/`` importable as the Python package
``lwfbrook90.jl_b200`` (a dot cannot appear in a plain package directory name)."""
import importlib.util as _u
import os as _os
import sys as _sys

_d = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "lwfbrook90.jl_b200")
if "lwfbrook90.jl_b200" not in _sys.modules:
    _spec = _u.spec_from_file_location("lwfbrook90.jl_b200", _os.path.join(_d, "__init__.py"),
                                       submodule_search_locations=[_d])
    jl_b200 = _u.module_from_spec(_spec)
    _sys.modules["lwfbrook90.jl_b200"] = jl_b200
    _spec.loader.exec_module(jl_b200)
else:
    jl_b200 = _sys.modules["lwfbrook90.jl_b200"]
