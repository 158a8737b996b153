"""`ptdeco_b200.utils`: the names the reference exposes under `ptdeco.utils` (configuration
schema, glue helpers, rank-search metrics), re-exported from the three submodules that hold them,
plus `relieve_gpu_memory_pressure`."""
from . import common as common
from . import losses_primitives as losses_primitives
from . import modconfig as modconfig

__all__: list[str] = []
for _mod in (modconfig, common, losses_primitives):
    for _name in _mod.__all__:
        globals()[_name] = getattr(_mod, _name)
        __all__.append(_name)
del _mod, _name
