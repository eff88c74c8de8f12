"""The data formats either side of the render path and a loader built for a renderer that takes
milliseconds per frame.

Formats read (written by the reference's data_preprocessing/fit_video.py:336-339,353-418, consumed by
dataloader/dataloader.py:38-71,146-230):

  split file (JSON)  { "img_res": int, "mutiview_intr_ls": [[fx, fy, cx, cy], ...] (focal lengths in pixels, principal point as
                       a fraction of the image size), optional "bg_path": [...],
                       "frames": [ { "fidx": int, "inst_dir": str, "head_transformation": [[4x4]],
                                     "mutiview_info_ls": [ { "view_name": str, "transform_matrix": [[4x4]] (camera to world),
                                                             "transform_matrix_ori": [[4x4]], "file_path": str, "mask_path": str,
                                                             optional "cam_K": [fx, fy, cx, cy] }, ... ] }, ... ] }
  condition images   <inst_dir>/ortho_{front,left,right}_{render,normal}_256_baseGama.png  (8-bit RGB)
  checkpoints        stage one: {"iter", "optimizer_state_dict", "loss", "psnr", "trainer_state_dict"}   (train_avatar.py:296-307)
                     stage two: {"iter", "nerf_optimizer", "g_optim", "d_optim", "nerf_render", "g", "d", "g_ema",
                                 "latent_codes"}                                                        (train_avatarHD.py:347-358)

What is different from the reference's loader: a frame is described to the renderer by an 18-float camera block (the rays are
generated inside the render kernel; the reference builds and uploads a [H*W, 11] float32 ray tensor per frame, 12.6 MB at
512 x 512), the six condition PNGs travel to the device as uint8 and are converted there (hav_make_render_cond), and decoded
condition tensors are cached ON THE DEVICE per instance directory (a 256 x 256 x 7 float32 triple is 5.5 MB: thousands of frames
fit in a corner of 180 GB), so a replayed or revisited frame costs no host work at all.  PNG decoding itself is cv2's, as in
the reference (a library call on the host, off the per-frame path once cached)."""
import ctypes as C
import json
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib

VIEWS = ("front", "left", "right")


def _imread_rgb(path):
    import cv2

    img = cv2.imread(path)
    if img is None:
        raise FileNotFoundError(path)
    return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


def read_cond_uint8(inst_dir, res):
    """The six PNGs of one frame as two uint8 arrays [3, res, res, 3] (render, normal), resized with cv2.INTER_LINEAR when
    their size differs from `res` like dataloader.py:219-225."""
    import cv2

    render, normal = [], []
    for v in VIEWS:
        for kind, dst in (("normal", normal), ("render", render)):
            img = _imread_rgb(os.path.join(inst_dir, "ortho_%s_%s_256_baseGama.png" % (v, kind)))
            if img.shape[0] != res:
                img = cv2.resize(img, dsize=(res, res), interpolation=cv2.INTER_LINEAR)
            dst.append(np.ascontiguousarray(img))
    return np.stack(render), np.stack(normal)


def make_render_cond(render_u8, normal_u8, device="cuda"):
    """uint8 [n, H, W, 3] pairs (host arrays or device tensors) -> float32 device tensor [n, 7, H, W] (hav_make_render_cond)."""
    L = _lib.lib()
    dev = torch.device(device)
    r = torch.as_tensor(render_u8).to(dev, non_blocking=True).contiguous()
    m = torch.as_tensor(normal_u8).to(dev, non_blocking=True).contiguous()
    if r.dtype != torch.uint8 or m.dtype != torch.uint8 or r.shape != m.shape or r.dim() != 4 or r.shape[-1] != 3:
        raise _lib.HavError("render / normal must be uint8 [n, H, W, 3] of equal shape")
    n, H, W = int(r.shape[0]), int(r.shape[1]), int(r.shape[2])
    out = torch.empty((n, 7, H, W), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.hav_make_render_cond(C.c_void_p(out.data_ptr()), C.c_void_p(r.data_ptr()), C.c_void_p(m.data_ptr()), n, H * W,
                                          C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "hav_make_render_cond")
    return out


def inv_head_T(head_transformation):
    """dataloader.py:206,215-216: [4,3] = [R^-1 ; -t] from the frame's 4x4 head transformation."""
    ht = np.asarray(head_transformation).astype(np.float32)[:3]
    rotation, translation = ht.T[:3, :3], ht.T[-1:]
    return np.concatenate([np.linalg.inv(rotation), -translation], 0).astype(np.float32)


class FrameDataset:
    """The (frame, view) list of a split file, in the reference's order (dataloader.py:50-71: every view except view_name '8',
    sorted by fidx), with per-item accessors that return what the renderer needs."""

    def __init__(self, split_file, down_sample=1.0, near=-1.6, far=1.0, length=1.0, cond_res=256, root=None):
        with open(split_file) as f:
            meta = json.load(f)
        self.root = os.path.dirname(os.path.abspath(split_file)) if root is None else root
        self.img_w = self.img_h = int(meta["img_res"])
        self.intrinsics = np.asarray(meta["mutiview_intr_ls"], dtype=np.float32)
        self.down_sample = float(down_sample)
        if self.down_sample < 1:                                                   # dataloader.py:52-54
            self.intrinsics[:, :2] = self.intrinsics[:, :2] * self.down_sample
            self.img_w = self.img_h = int(self.img_w * self.down_sample)
        self.near, self.far, self.length, self.cond_res = float(near), float(far), float(length), int(cond_res)
        self.bg_paths = meta.get("bg_path")
        frames = []
        for fr in meta["frames"]:
            for vidx, view in enumerate(fr["mutiview_info_ls"]):
                if view["view_name"] == "8":
                    continue
                frames.append((fr, vidx))
        frames.sort(key=lambda fv: fv[0]["fidx"])                                 # stable, like list.sort in the reference
        self.frames = frames

    def __len__(self):
        return len(self.frames)

    def path(self, p):
        return p if os.path.isabs(p) else os.path.join(self.root, p)

    def fidx(self, idx):
        return int(self.frames[idx][0]["fidx"])

    def camera(self, idx):
        """(intr [4], c2w [3,4], near, far) of item idx: dataloader.py:140-152,174-177."""
        fr, vidx = self.frames[idx]
        view = fr["mutiview_info_ls"][vidx]
        pose = np.asarray(view["transform_matrix"], dtype=np.float32)
        if "cam_K" in view:
            intr = np.asarray(view["cam_K"], dtype=np.float32).copy()
            if self.down_sample < 1:
                intr[:2] = intr[:2] * self.down_sample
        else:
            intr = self.intrinsics[vidx]
        t_ori = np.asarray(view["transform_matrix_ori"], dtype=np.float32)[:3, -1]
        dist = np.float32(np.linalg.norm(t_ori))
        return intr, pose[:3, :4], np.float32(dist + np.float32(self.near * self.length)), np.float32(dist + np.float32(self.far * self.length))

    def camera_block(self, indices, device="cuda"):
        """[B,18] device tensor for render_rays(camera=...)."""
        from .render import camera_block

        cams = [self.camera(i) for i in indices]
        return camera_block(np.stack([c[0] for c in cams]), np.stack([c[1] for c in cams]),
                            np.array([c[2] for c in cams], dtype=np.float32), np.array([c[3] for c in cams], dtype=np.float32), device=device)

    def inv_head_T(self, idx):
        return inv_head_T(self.frames[idx][0]["head_transformation"])

    def inst_dir(self, idx):
        return self.path(self.frames[idx][0]["inst_dir"])

    def cond_uint8(self, idx):
        return read_cond_uint8(self.inst_dir(idx), self.cond_res)

    def image_and_mask(self, idx, mask_thresh=127.5):
        """Ground-truth colour [H,W,3] float32 composited over white and mask [H,W] (dataloader.py:154-161,183-190), for the
        training / validation targets."""
        import cv2

        view = self.frames[idx][0]["mutiview_info_ls"][self.frames[idx][1]]
        mask = _imread_rgb(self.path(view["mask_path"]))
        img = _imread_rgb(self.path(view["file_path"]))
        if self.down_sample < 1:
            mask = cv2.resize(mask, dsize=(0, 0), fx=self.down_sample, fy=self.down_sample, interpolation=cv2.INTER_AREA)
            img = cv2.resize(img, dsize=(0, 0), fx=self.down_sample, fy=self.down_sample, interpolation=cv2.INTER_AREA)
        m = (mask[:, :, 0] > mask_thresh).astype(np.float32)
        rgb = (np.array(img) / 255.0).astype(np.float32)
        return rgb * m[..., None] + np.float32(1.0) * (np.float32(1.0) - m[..., None]), m


class CondCache:
    """Device-resident cache of the three condition tensors of a frame ([3,7,res,res] float32: front, left, right), keyed by
    instance directory, LRU-bounded by bytes.  A miss decodes the six PNGs on the host (cv2), uploads 6 x res^2 x 3 bytes and
    converts on the device; a hit is a dictionary lookup."""

    def __init__(self, device="cuda", max_bytes=8 << 30):
        self.device, self.max_bytes = torch.device(device), int(max_bytes)
        self._items, self._bytes = OrderedDict(), 0
        self.hits = self.misses = 0

    def get(self, inst_dir, res):
        key = (inst_dir, res)
        t = self._items.get(key)
        if t is not None:
            self._items.move_to_end(key)
            self.hits += 1
            return t
        self.misses += 1
        render, normal = read_cond_uint8(inst_dir, res)
        t = make_render_cond(render, normal, self.device)
        self._items[key] = t
        self._bytes += t.numel() * 4
        while self._bytes > self.max_bytes and len(self._items) > 1:
            _, old = self._items.popitem(last=False)
            self._bytes -= old.numel() * 4
        return t


class FrameLoader:
    """Batches of validation / reenactment frames for havatar_b200.trainer.Trainer.forward (mode 'validation' / 'test' of
    dataloader.py:193-216): `batch(indices)` returns the dict of device tensors the Trainer takes, with `camera` / `img_hw`
    instead of a ray tensor.  Backgrounds are white (`white_bg=True`, the only mode the entry scripts use)."""

    def __init__(self, dataset, device="cuda", cache_bytes=8 << 30):
        self.ds, self.device = dataset, torch.device(device)
        self.cache = CondCache(device, cache_bytes)
        self._bg = {}

    def batch(self, indices):
        ds, dev = self.ds, self.device
        conds = torch.stack([self.cache.get(ds.inst_dir(i), ds.cond_res) for i in indices])        # [B,3,7,res,res]
        R = ds.img_h * ds.img_w
        key = (len(indices), R)
        if key not in self._bg:
            self._bg[key] = torch.ones((len(indices), R, 3), dtype=torch.float32, device=dev)
        return {"fidx": [ds.fidx(i) for i in indices], "camera": ds.camera_block(indices, dev), "img_hw": (ds.img_h, ds.img_w),
                "background_prior": self._bg[key],
                "inv_head_T": torch.from_numpy(np.stack([ds.inv_head_T(i) for i in indices])).to(dev),
                "front_render_cond": conds[:, 0], "left_render_cond": conds[:, 1], "right_render_cond": conds[:, 2]}


# ---- checkpoint dictionaries ------------------------------------------------------------------------------------------------
def stage_one_checkpoint(it, trainer, optimizer, loss=None, psnr=None):
    """train_avatar.py:296-307."""
    return {"iter": it, "optimizer_state_dict": optimizer.state_dict(), "loss": loss, "psnr": psnr, "trainer_state_dict": trainer.state_dict()}


def stage_two_checkpoint(it, nerf_render, generator, discriminator, g_ema, nerf_optimizer, g_optim, d_optim):
    """train_avatarHD.py:347-358."""
    return {"iter": it, "nerf_optimizer": nerf_optimizer.state_dict(), "g_optim": g_optim.state_dict(), "d_optim": d_optim.state_dict(),
            "nerf_render": nerf_render.state_dict(), "g": generator.state_dict(), "d": discriminator.state_dict(),
            "g_ema": g_ema.state_dict(), "latent_codes": nerf_render.latent_codes.data}


def load_partial_state_dict(model, state, except_keys=()):
    """utils/training_util.py:124-139 (full_name=False): every entry of `state` whose name does not start with one of
    `except_keys` replaces the model's, then a strict load_state_dict."""
    own = model.state_dict()
    own.update({k: v for k, v in state.items() if not any(k.startswith(e) for e in except_keys)})
    model.load_state_dict(own)


def load_reenactment_checkpoint(ckpt, nerf_render, img_trans):
    """avatarHD_reenactment.py:138-146: stage-two checkpoint -> inference networks."""
    load_partial_state_dict(nerf_render, ckpt["nerf_render"], except_keys=["latent_codes"])
    nerf_render.latent_codes = ckpt["latent_codes"].to(next(nerf_render.parameters()).device)      # :142 (built with 0 codes)
    img_trans.load_state_dict(ckpt["g_ema"])
    nerf_render.headpose_skin_net.fix_canonical_W()
    return nerf_render.eval(), img_trans.eval()
