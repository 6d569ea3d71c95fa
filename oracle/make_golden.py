#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference in the build container.

TEST INFRASTRUCTURE.  Run here (``python oracle/make_golden.py``); needs /root/reference, which
does not exist on the GPU box -- the committed .npz files are what travels.

The reference's executable spec ``similarity_map(ssl_mode='pytorch')``
(GAN-Based-SR/basicsr/losses/loss_util.py:165-229) is imported unmodified.  A plain import fails
on a GPU-less box because loss_util.py:9 pulls in the JIT CUDA wrapper, which calls sys.exit()
without CUDA (similaritywrapper.py:11-13), and ``basicsr/__init__`` needs lmdb; so the parent
packages are pre-seeded with empty stubs and the file is loaded by path.
The L1 is ``loss_weight * F.l1_loss`` (basic_loss.py:14-16,66) and the KL is the expression of
basic_loss.py:280; gradients come from the reference's own autograd graph.
The mask fixture runs the statements of scripts/data_preparation/generate_mask.py:22-31
(PIL convert("L") -> cv2.Laplacian(CV_8U) -> > 20) on a crop of the reference's baboon.png.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
REF = os.environ.get("SSL_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")


def load_reference_similarity_map():
    for name in ("basicsr", "basicsr.losses", "basicsr.losses.similarity"):
        mod = types.ModuleType(name)
        mod.__path__ = []
        sys.modules.setdefault(name, mod)
    stub = types.ModuleType("basicsr.losses.similarity.similaritywrapper")
    stub.compute_similarity = None
    sys.modules["basicsr.losses.similarity.similaritywrapper"] = stub
    path = os.path.join(REF, "GAN-Based-SR", "basicsr", "losses", "loss_util.py")
    spec = importlib.util.spec_from_file_location("basicsr.losses.loss_util", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.similarity_map


def run_reference(similarity_map, sr, gt, mask, ks, kw, sigma, gen, dtype, loss_weight=1.0, kl_weight=1.0):
    """Reference training-step block (realesrganssl_model.py:378-426) for a batch, in `dtype`."""
    sr = sr.to(dtype).clone().requires_grad_(True)
    gt = gt.to(dtype)
    mask = mask.to(dtype)
    rows_sr, rows_gt = [], []
    for i in range(sr.shape[0]):
        m = mask[i, :].unsqueeze(0)
        if m.sum() == 0:
            continue
        s = similarity_map(img=sr[i, :].unsqueeze(0).clone(), mask=m.clone(), ssl_mode="pytorch",
                           kernel_size_search=ks, generalization=gen, kernel_size_window=kw, sigma=sigma).getitem()
        t = similarity_map(img=gt[i, :].unsqueeze(0).clone(), mask=m.clone(), ssl_mode="pytorch",
                           kernel_size_search=ks, generalization=gen, kernel_size_window=kw, sigma=sigma).getitem()
        rows_sr.append(s)
        rows_gt.append(t)
    s = torch.cat(rows_sr, dim=1)
    t = torch.cat(rows_gt, dim=1)
    l1 = loss_weight * F.l1_loss(s, t, reduction="none").mean()
    kl = kl_weight * F.kl_div(torch.clamp(s, min=1e-10).log(), torch.clamp(t, min=1e-10), reduction="mean")
    g_l1, = torch.autograd.grad(l1, sr, retain_graph=True)
    g_kl, = torch.autograd.grad(kl, sr)
    return dict(rows_sr=s[0].detach().numpy(), rows_gt=t[0].detach().numpy(), l1=float(l1), kl=float(kl),
                grad_l1=g_l1.numpy(), grad_kl=g_kl.numpy())


def save_case(name, similarity_map, sr, gt, mask, ks, kw, sigma=0.004, gen=True):
    out = dict(sr=sr.numpy(), gt=gt.numpy(), mask=mask.numpy(), ks=ks, kw=kw, sigma=sigma, gen=int(gen))
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        res = run_reference(similarity_map, sr, gt, mask, ks, kw, sigma, gen, dt)
        for k, v in res.items():
            out[f"{k}_{tag}"] = v
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: rows {out['rows_sr_f32'].shape}  l1(f32)={out['l1_f32']:.6e} l1(f64)={out['l1_f64']:.6e} "
          f"kl(f64)={out['kl_f64']:.6e}  max|g|={np.abs(out['grad_l1_f64']).max():.3e}  "
          f"{os.path.getsize(path) / 1e3:.0f} kB")


def mask_fixture():
    import cv2
    from PIL import Image
    img = Image.open(os.path.join(REF, "GAN-Based-SR", "test_scripts", "data", "baboon.png"))
    rgb = np.array(img.convert("RGB"))[100:196, 60:188]  # 96 x 128 crop, uint8 HWC
    gray = np.array(Image.fromarray(rgb).convert("L"))
    lap = cv2.Laplacian(gray, cv2.CV_8U)
    mask = np.zeros(gray.shape, dtype="int")
    mask[lap > 20.0] = 1
    path = os.path.join(OUT, "laplacian_mask.npz")
    np.savez_compressed(path, rgb=np.ascontiguousarray(rgb.transpose(2, 0, 1)), gray=gray, lap=lap,
                        mask=mask.astype(np.float32), threshold=20.0)
    print(f"laplacian_mask: density {mask.mean():.3f}  {os.path.getsize(path) / 1e3:.0f} kB")


def main():
    from ssl_b200 import synth
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    similarity_map = load_reference_similarity_map()

    # config A (BASELINE.json configs[0]): single 64x64 crop, k_s=11, k_w=5
    sr, gt, mask = synth.make_case(1, 64, 64, seed=0, density=0.03)
    save_case("configA_seed0", similarity_map, sr, gt, mask, 11, 5)
    sr, gt, mask = synth.make_case(2, 64, 64, seed=3, density=0.02)
    mask[1] = 0  # image with an empty mask is skipped by the caller
    mask[1, 0, 31, 17] = 1  # ... and then given a single interior pixel
    save_case("configA_batch2_single_pixel", similarity_map, sr, gt, mask, 11, 5)
    sr, gt, mask = synth.make_case(1, 64, 64, seed=4, density=0.02)
    save_case("configA_no_generalization", similarity_map, sr, gt, mask, 11, 5, gen=False)
    sr, gt, mask = synth.make_case(1, 64, 64, seed=5, density=0.01)
    save_case("configA_mask3ch", similarity_map, sr, gt, mask.repeat(1, 3, 1, 1), 11, 5)
    # the headline kernel sizes (k_s=25, k_w=9) on a crop small enough for the torch oracle
    sr, gt, mask = synth.make_case(1, 48, 56, seed=6, density=0.012)
    save_case("k25w9_48x56", similarity_map, sr, gt, mask, 25, 9)
    # k_w == k_s edge of the K <= P contract, non-square, tiny
    sr, gt, mask = synth.make_case(1, 20, 24, seed=7, density=0.05)
    save_case("k7w7_20x24", similarity_map, sr, gt, mask, 7, 7, sigma=0.05)
    mask_fixture()


if __name__ == "__main__":
    main()
