"""Run the reference's own plugin files on this package's ``Attack`` base.

The reference's ~120 attack plugins reach the base class and the helpers only through the relative imports
``from ..attack import Attack`` / ``from ..utils import *`` (transferattack/gradient/mifgsm.py:3-4 and every other
plugin). ``adopt_reference_plugins`` builds a package whose ``attack`` and ``utils`` sub-modules are THIS package's
modules and whose remaining sub-packages (``gradient``, ``input_transformation``, ``ensemble``, ...) are imported
from a reference checkout on disk, unmodified. Every hook those plugins call (``get_momentum``, ``update_delta``,
``init_delta``, ``get_grad`` ...) then lands in the sm_100a kernels.

    import transferattack_b200.compat as compat
    ta = compat.adopt_reference_plugins('/path/to/TransferAttack')      # -> module with attack_zoo / load_attack_class
    attacker = ta.load_attack_class('gra')(model_name='resnet50')
"""
import importlib.util
import os
import sys

from . import attack as _attack
from . import utils as _utils


def adopt_reference_plugins(reference_root, package_name="transferattack"):
    root = os.path.join(reference_root, "transferattack")
    init = os.path.join(root, "__init__.py")
    if not os.path.isfile(init):
        raise FileNotFoundError("no transferattack/__init__.py under {}".format(reference_root))
    if package_name in sys.modules:
        mod = sys.modules[package_name]
        if getattr(mod, "__ta_b200_adopted__", None) == root:
            return mod
        raise RuntimeError("a module named {!r} is already imported; pass another package_name".format(package_name))
    spec = importlib.util.spec_from_file_location(package_name, init, submodule_search_locations=[root])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[package_name] = mod
    # seed the two modules the plugins import relatively BEFORE anything of the reference is executed
    sys.modules[package_name + ".attack"] = _attack
    sys.modules[package_name + ".utils"] = _utils
    spec.loader.exec_module(mod)          # the reference's registry: attack_zoo + lazy load_attack_class
    mod.attack, mod.utils = _attack, _utils
    mod.__ta_b200_adopted__ = root
    return mod
