"""Registry with the API of basicsr.utils.registry.Registry (Car_Plate-Restoration/basicsr/utils/registry.py:4-82):
`register()` as decorator or call, `get(name)` (KeyError when missing), `__contains__`, `keys()`, iteration, and an
assertion on duplicate names.  When the reference's own registry module is already imported in this process the B200
arch registers into that one (so `build_network(opt)` finds it); otherwise this stand-alone twin is used.
"""
import sys


class Registry:
    """name -> class table.  `_obj_map` is kept as the storage attribute because callers of the reference's registry
    (and register_into(override=True) here) reach into it directly."""

    def __init__(self, name):
        self._name, self._obj_map = name, {}

    def _do_register(self, name, obj):
        if name in self._obj_map:       # the reference asserts on duplicates (registry.py:38-41); same exception type
            raise AssertionError(f"'{name}' is already present in the '{self._name}' registry")
        self._obj_map[name] = obj

    def register(self, obj=None):
        """@REGISTRY.register() on a class / function, or REGISTRY.register(obj)."""
        def add(target):
            self._do_register(target.__name__, target)
            return target
        return add if obj is None else (add(obj) and None)

    def get(self, name):
        try:
            return self._obj_map[name]
        except KeyError:
            raise KeyError(f"'{name}' is not in the '{self._name}' registry (known: {sorted(self._obj_map)})") from None

    def keys(self):
        return self._obj_map.keys()

    def __contains__(self, name):
        return name in self._obj_map

    def __iter__(self):
        yield from self._obj_map.items()


_ref = sys.modules.get('basicsr.utils.registry')
if _ref is not None and hasattr(_ref, 'ARCH_REGISTRY'):
    ARCH_REGISTRY = _ref.ARCH_REGISTRY
    USING_BASICSR_REGISTRY = True
else:
    ARCH_REGISTRY = Registry('arch')
    USING_BASICSR_REGISTRY = False


def build_network(opt):
    """basicsr.archs.build_network (archs/__init__.py:19-25): pops `type`, instantiates with the remaining keys."""
    opt = dict(opt)
    network_type = opt.pop('type')
    return ARCH_REGISTRY.get(network_type)(**opt)
