"""transferattack_b200 — B200-native engine for TransferAttack's iterative hot loop.

Same registry surface as the reference package (``attack_zoo``, ``load_attack_class``; transferattack/__init__.py:3-160)
for the attacks on the accelerated path; every other reference plugin runs unchanged on this base class through
``transferattack_b200.compat.adopt_reference_plugins``.
"""
import importlib

from . import utils, attack   # noqa: F401  (reference-style `pkg.utils.wrap_model` / `pkg.attack.Attack` access)

attack_zoo = {
    # gradient
    'fgsm': ('.gradient.fgsm', 'FGSM'),
    'ifgsm': ('.gradient.ifgsm', 'IFGSM'),
    'mifgsm': ('.gradient.mifgsm', 'MIFGSM'),
    'nifgsm': ('.gradient.nifgsm', 'NIFGSM'),
    'vmifgsm': ('.gradient.vmifgsm', 'VMIFGSM'),
    'vnifgsm': ('.gradient.vnifgsm', 'VNIFGSM'),
    'emifgsm': ('.gradient.emifgsm', 'EMIFGSM'),
    # input transformation
    'dim': ('.input_transformation.dim', 'DIM'),
    'tim': ('.input_transformation.tim', 'TIM'),
    'sim': ('.input_transformation.sim', 'SIM'),
    'admix': ('.input_transformation.admix', 'Admix'),
    'ditimi': ('.input_transformation.di_ti_mi', 'DITIMI'),
    # ensemble
    'ens': ('.ensemble.ens', 'ENS'),
}


def load_attack_class(attack_name):
    if attack_name not in attack_zoo:
        raise Exception('Unspported attack algorithm {}'.format(attack_name))
    module_path, class_name = attack_zoo[attack_name]
    module = importlib.import_module(module_path, __package__)
    return getattr(module, class_name)


__version__ = '0.1.0'
