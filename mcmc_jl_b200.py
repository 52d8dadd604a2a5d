"""Import shim: the product package lives in the directory `mcmc.jl_b200/` (the layout this repo is
required to have), whose name is not a Python identifier.  `import mcmc_jl_b200` loads that package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mcmc.jl_b200")
_spec = importlib.util.spec_from_file_location("mcmc_jl_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mcmc_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
