"""Install the UNMODIFIED reference into baseline/_ref/ (git-ignored; it travels to the GPU box with the repo snapshot) so that
bench.py can time the reference's own GPU path -- the >= 10x denominator BASELINE.md section 4 / SURVEY.md section 8(d) name.

    python baseline/install_ref.py            # build container only: needs /root/reference, nvcc, torch headers

What it does (nothing under baseline/_ref/ is tracked; no reference source enters the repository):
  1. copies the reference's Python packages the render path imports (model/, utils/, config/, dataloader/data_util.py) from
     /root/reference to baseline/_ref/havatar/ byte for byte (the reference has no setup.py, so `pip install` has nothing to
     build: README.md:16-21 installs only the two extensions);
  2. builds the reference's own CUDA extensions from model/op (model/op/setup.py:5-15: `upfirdn2d` = upfirdn2d.cpp +
     upfirdn2d_kernel.cu, `fused` = fused_bias_act.cpp + fused_bias_act_kernel.cu) for sm_100a with
     torch.utils.cpp_extension into baseline/_ref/ext/{upfirdn2d,fused}.so, the bare-name modules model/op/*.py import.
bench.py --reference-gpu (and the `reference_gpu` key of the default run) puts both directories on sys.path, applies the two
environment shims the reference needs on a box without matplotlib (utils/training_util.py:6) and imports
model.nerf_trainer.Trainer / model.styleUnet.SWGAN_unet exactly as train_avatarHD.py does."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HAV_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def copy_python():
    dst = os.path.join(DST, "havatar")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    for pkg in ("model", "utils", "config"):
        shutil.copytree(os.path.join(REF, pkg), os.path.join(dst, pkg), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    os.makedirs(os.path.join(dst, "dataloader"))
    shutil.copy2(os.path.join(REF, "dataloader", "data_util.py"), os.path.join(dst, "dataloader", "data_util.py"))
    shutil.copy2(os.path.join(REF, "LICENSE"), os.path.join(dst, "LICENSE"))
    return dst


def build_ext():
    """model/op/setup.py's two CUDAExtensions, built out of tree (the reference tree is read-only)."""
    ext = os.path.join(DST, "ext")
    work = os.path.join(DST, "build")
    os.makedirs(ext, exist_ok=True)
    if os.path.isdir(work):
        shutil.rmtree(work)
    shutil.copytree(os.path.join(REF, "model", "op"), work)
    env = dict(os.environ, TORCH_CUDA_ARCH_LIST="10.0a", MAX_JOBS=str(os.cpu_count() or 4))
    r = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=work, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-4000:] + r.stderr[-4000:])
        raise SystemExit("building the reference's model/op failed")
    built = [f for f in os.listdir(work) if f.endswith(".so")]
    for f in built:
        shutil.copy2(os.path.join(work, f), os.path.join(ext, f))
    shutil.rmtree(work)
    return sorted(built)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not found at %s" % REF)
    os.makedirs(DST, exist_ok=True)
    print("python packages ->", copy_python())
    print("extensions ->", build_ext())
