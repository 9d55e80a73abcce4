"""hcflow_b200 -- B200 (sm_100a) flow-step engine behind HCFlow's own module API.

    import hcflow_b200
    hcflow_b200.install()      # before the reference's networks.define_G(opt, step) runs

makes the reference's factory (codes/models/networks.py:9-41) resolve
``which_model_G: HCFlowNet_SR`` / ``HCFlowNet_Rescaling`` to the CUDA-backed classes of
``hcflow_b200.arch``; everything else of the reference (YAML options, model wrappers,
test_HCFlow.py, checkpoints) is used unchanged.  See INTEGRATION.md.
"""
import sys
import types

__version__ = "0.1.0"

_ARCH_MODULES = {
    "models.modules.HCFlowNet_SR_arch": "HCFlowNet_SR",
    "models.modules.HCFlowNet_Rescaling_arch": "HCFlowNet_Rescaling",
}


def install():
    """Register the drop-in arch modules under the import names the reference's
    ``find_model_using_name`` uses (importlib returns sys.modules entries first)."""
    from . import arch
    for modname, clsname in _ARCH_MODULES.items():
        m = types.ModuleType(modname)
        m.__doc__ = "hcflow_b200 drop-in for the reference module of the same name"
        setattr(m, clsname, getattr(arch, clsname))
        sys.modules[modname] = m
    return sorted(_ARCH_MODULES)


def uninstall():
    for modname in _ARCH_MODULES:
        sys.modules.pop(modname, None)
