"""Import shims that let the UNMODIFIED reference modules run on CPU in the build
container.  Only usable where /root/reference exists; nothing on the
GPU box may call this.  No reference code is copied -- the shims only satisfy imports.

  D12: ccdm/ddpm/__init__.py star-imports the trainer (ignite, medpy ...) -> pre-register
       an empty ``ddpm`` package whose __path__ points at the reference directory.
  D6 : latentdiffusion imports a missing ``models.util`` -> stub exposing
       ``instantiate_from_config``.
  D7 : UNetModel lazily imports omegaconf.listconfig.ListConfig -> stub.
  D8 : DDIMSampler.register_buffer force-moves tensors to "cuda" -> CPU subclass.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("GUIDEGEN_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "ccdm", "ddpm", "models"))


def _install():
    sys.dont_write_bytecode = True  # reference tree is read-only
    if "ddpm" not in sys.modules:
        pkg = types.ModuleType("ddpm")
        pkg.__path__ = [os.path.join(REF_ROOT, "ccdm", "ddpm")]
        sys.modules["ddpm"] = pkg
    if "models" not in sys.modules:
        m = types.ModuleType("models")
        m.__path__ = []
        u = types.ModuleType("models.util")

        def instantiate_from_config(config):  # pragma: no cover - never reached on the hot path
            raise RuntimeError("instantiate_from_config is not available in the oracle shim")

        u.instantiate_from_config = instantiate_from_config
        m.util = u
        sys.modules["models"] = m
        sys.modules["models.util"] = u
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")
        lc = types.ModuleType("omegaconf.listconfig")

        class ListConfig(list):
            pass

        lc.ListConfig = ListConfig
        oc.listconfig = lc
        sys.modules["omegaconf"] = oc
        sys.modules["omegaconf.listconfig"] = lc
    ld = os.path.join(REF_ROOT, "latentdiffusion")
    if ld not in sys.path:
        sys.path.insert(0, ld)


def ccdm():
    """-> namespace with build_model, DiffusionModel, DenoisingModel, OneHotCategoricalBCHW, UNetModel."""
    if not available():
        raise RuntimeError("reference tree not present")
    _install()
    ns = types.SimpleNamespace()
    b = importlib.import_module("ddpm.models.builder")
    dd = importlib.import_module("ddpm.models.diffusion_denoising")
    oh = importlib.import_module("ddpm.models.one_hot_categorical")
    un = importlib.import_module("ddpm.models.unet_openai.unet")
    ns.build_model = b.build_model
    ns.DiffusionModel = dd.DiffusionModel
    ns.DenoisingModel = dd.DenoisingModel
    ns.OneHotCategoricalBCHW = oh.OneHotCategoricalBCHW
    ns.UNetModel = un.UNetModel
    ns.unet_module = un
    return ns


def ldm():
    """-> namespace with UNetModel, SpatialTransformer, DDIMSamplerCPU, util."""
    if not available():
        raise RuntimeError("reference tree not present")
    _install()
    ns = types.SimpleNamespace()
    om = importlib.import_module("ldm.modules.diffusionmodules.openaimodel")
    at = importlib.import_module("ldm.modules.attention")
    ut = importlib.import_module("ldm.modules.diffusionmodules.util")
    dm = importlib.import_module("ldm.models.diffusion.ddim")

    class DDIMSamplerCPU(dm.DDIMSampler):
        def register_buffer(self, name, attr):  # D8: keep tensors where they are
            setattr(self, name, attr)

    pm = importlib.import_module("ldm.models.diffusion.plms")

    class PLMSSamplerCPU(pm.PLMSSampler):
        def register_buffer(self, name, attr):  # D8: keep tensors where they are
            setattr(self, name, attr)

    ns.PLMSSamplerCPU = PLMSSamplerCPU
    ns.plms_module = pm
    ns.UNetModel = om.UNetModel
    ns.SpatialTransformer = at.SpatialTransformer
    ns.CrossAttention = at.CrossAttention
    ns.BasicTransformerBlock = at.BasicTransformerBlock
    ns.DDIMSampler = dm.DDIMSampler
    ns.DDIMSamplerCPU = DDIMSamplerCPU
    ns.util = ut
    ns.ddim_module = dm
    return ns
