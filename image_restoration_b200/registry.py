"""Registry with the API of basicsr.utils.registry.Registry (Car_Plate-Restoration/basicsr/utils/registry.py:4-82):
`register()` as decorator or call, `get(name)` (KeyError when missing), `__contains__`, `keys()`, iteration, and an
assertion on duplicate names.  When the reference's own registry module is already imported in this process the B200
arch registers into that one (so `build_network(opt)` finds it); otherwise this stand-alone twin is used.
"""
import sys


class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj):
        assert name not in self._obj_map, (f"An object named '{name}' was already registered "
                                           f"in '{self._name}' registry!")
        self._obj_map[name] = obj

    def register(self, obj=None):
        if obj is None:
            def deco(func_or_class):
                self._do_register(func_or_class.__name__, func_or_class)
                return func_or_class
            return deco
        self._do_register(obj.__name__, obj)

    def get(self, name):
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return ret

    def __contains__(self, name):
        return name in self._obj_map

    def __iter__(self):
        return iter(self._obj_map.items())

    def keys(self):
        return self._obj_map.keys()


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
