"""transferattack_b200 — B200-native engine for TransferAttack's iterative hot loop.

Same registry surface as the reference package (``attack_zoo``, ``load_attack_class``; transferattack/__init__.py:3-160)
for the attacks on the accelerated path; every other reference plugin runs unchanged on this base class through
``transferattack_b200.compat.adopt_reference_plugins``.
"""
import importlib

from . import utils, attack   # noqa: F401  (reference-style `pkg.utils.wrap_model` / `pkg.attack.Attack` access)

def _zoo():
    """name → (relative module, class); same keys and classes as the reference registry for the accelerated attacks."""
    table = {
        "gradient": ["fgsm:FGSM", "ifgsm:IFGSM", "mifgsm:MIFGSM", "nifgsm:NIFGSM", "vmifgsm:VMIFGSM", "vnifgsm:VNIFGSM",
                     "emifgsm:EMIFGSM", "pifgsm:PIFGSM", "gra:GRA"],
        "input_transformation": ["dim:DIM", "tim:TIM", "sim:SIM", "admix:Admix", "di_ti_mi:DITIMI=ditimi", "di_ti_mi:SIDITIMI=siditimi", "ssm:SSM"],
        "ensemble": ["ens:ENS", "adaea:AdaEA"],
    }
    zoo = {}
    for package, entries in table.items():
        for entry in entries:
            spec, _, alias = entry.partition("=")
            module, cls = spec.split(":")
            zoo[alias or module] = (".%s.%s" % (package, module), cls)
    return zoo


attack_zoo = _zoo()


def load_attack_class(attack_name):
    if attack_name not in attack_zoo:
        raise Exception('Unspported attack algorithm {}'.format(attack_name))
    module_path, class_name = attack_zoo[attack_name]
    module = importlib.import_module(module_path, __package__)
    return getattr(module, class_name)


__version__ = '0.1.0'
