"""Import shim that makes the UNMODIFIED reference (/root/reference) importable on CPU.

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py in the build container to mint golden
vectors; nothing on the product path, in the gpu tests, in smoke() or in bench.py may import this
module (the reference tree does not exist on the GPU box).

The four shims:
  1. empty stub modules `fused` / `upfirdn2d` (model/op/fused_act.py:20, model/op/upfirdn2d.py:19
     import them by bare name; on CPU tensors the wrappers branch to their pure-torch fallbacks
     at fused_act.py:108 and upfirdn2d.py:163 so the stubs are never called);
  2. stub `matplotlib` (utils/training_util.py:6 imports it, the render path never uses it);
  3. force get_embedder(device='cpu') (model/network/embedder.py:99 defaults to 'cuda');
  4. Tensor.cuda -> identity for StyleGAN_zxc.make_noise (model/styleUnet.py:748-751).
"""
import sys
import types

REF_ROOT = "/root/reference"
_installed = False


def install():
    global _installed
    if _installed:
        return
    import torch

    for name in ("fused", "upfirdn2d"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    import model.network.embedder as emb

    _orig = emb.get_embedder

    def _cpu_get_embedder(multires, i=0, input_dims=3, include_input=True, device="cpu"):
        return _orig(multires, i=i, input_dims=input_dims, include_input=include_input, device="cpu")

    emb.get_embedder = _cpu_get_embedder
    torch.Tensor.cuda = lambda self, *a, **k: self
    _installed = True


def load_cfg(name="singleview_512_base.yml"):
    install()
    import yaml
    from utils.cfgnode import CfgNode

    with open(f"{REF_ROOT}/config/{name}") as f:
        return CfgNode(yaml.load(f, Loader=yaml.FullLoader))
