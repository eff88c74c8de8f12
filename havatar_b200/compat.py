"""`model.*` namespace shim: lets the reference's entry scripts run UNMODIFIED on the B200 kernels.

train_avatar.py:20, train_avatarHD.py:19-20,27 and avatarHD_reenactment.py:6,14 import

    from model.nerf_trainer import Trainer
    from model.styleUnet import SWGAN_unet, Discriminator

and utils/styleUnet_util.py:7 imports `from model.op import conv2d_gradfix`.  `install()` registers modules under exactly those
names in sys.modules -- ahead of the reference's own `model/` package on sys.path -- whose attributes are the havatar_b200 drop-ins:

    import havatar_b200.compat as compat; compat.install()      # first lines of the entry script (or sitecustomize / -c)
    # ... the rest of train_avatar.py / train_avatarHD.py / avatarHD_reenactment.py unchanged

Everything else the scripts import (dataloader.*, utils.*) still comes from the reference tree.  `uninstall()` restores the
previous sys.modules entries."""
import sys
import types

_NAMES = ("model", "model.nerf_trainer", "model.styleUnet", "model.nerf_model", "model.Skinning_Field", "model.op",
          "model.op.fused_act", "model.op.upfirdn2d", "model.op.conv2d_gradfix", "fused", "upfirdn2d")
_saved = None


def _module(name, doc, **attrs):
    m = types.ModuleType(name, doc)
    m.__dict__.update(attrs)
    return m


def install():
    """Register the shim.  Idempotent.  Returns the list of module names it provides."""
    global _saved
    from . import op, styleunet, trainer
    from .op import conv2d_gradfix, fused, fused_act, upfirdn2d as upfirdn2d_mod, upfirdn2d_op

    if _saved is None:
        _saved = {n: sys.modules.get(n) for n in _NAMES}
    pkg = _module("model", "havatar_b200 stand-in for the reference's model/ package")
    pkg.__path__ = []                                   # a package: `import model.x` consults sys.modules first
    mods = {
        "model": pkg,
        "model.nerf_trainer": _module("model.nerf_trainer", "model/nerf_trainer.py -> havatar_b200.trainer", Trainer=trainer.Trainer),
        "model.styleUnet": _module(
            "model.styleUnet", "model/styleUnet.py -> havatar_b200.styleunet",
            **{k: getattr(styleunet, k) for k in ("SWGAN_unet", "StyleGAN_zxc", "Discriminator", "ModulatedConv2d", "StyledConv", "ToRGB",
                                                  "ConvLayer", "ConvBlock", "FromRGB", "EqualLinear", "EqualConv2d", "Blur", "Upsample",
                                                  "Downsample", "HaarTransform", "InverseHaarTransform", "ConstantInput", "PixelNorm",
                                                  "NoiseInjection") if hasattr(styleunet, k)}),
        "model.nerf_model": _module("model.nerf_model", "model/nerf_model.py -> havatar_b200.trainer.PlaneNeRF",
                                    ConditionalTriplaneNeRFModel_multiRender_split_view=trainer.PlaneNeRF),
        "model.Skinning_Field": _module("model.Skinning_Field", "model/Skinning_Field.py -> havatar_b200.trainer.SkinningField",
                                        Deformation_Field_new=trainer.SkinningField),
        "model.op": op, "model.op.fused_act": fused_act, "model.op.upfirdn2d": upfirdn2d_mod, "model.op.conv2d_gradfix": conv2d_gradfix,
        "fused": fused, "upfirdn2d": upfirdn2d_op,
    }
    for name, m in mods.items():
        sys.modules[name] = m
    for name in ("nerf_trainer", "styleUnet", "nerf_model", "Skinning_Field", "op"):
        setattr(pkg, name, mods["model." + name])
    return list(mods)


def uninstall():
    global _saved
    if _saved is None:
        return
    for n, m in _saved.items():
        if m is None:
            sys.modules.pop(n, None)
        else:
            sys.modules[n] = m
    _saved = None
