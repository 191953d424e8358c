"""ORACLE test infrastructure: freeze tests/golden/ref_img_eagle.npz from the reference's UNMODIFIED
eagle/Dataloader/IMG_Eagle.EagleDataset (development container only, needs /root/reference):

    python -m oracle.make_golden_img_eagle

A seeded synthetic `states.npy` (600 frames of a 16 x 32 x 4 grid) + `pixel_type.npy` under <tmp>/ds_img/<example>/<sim>/, the
split file under <tmp>/Splits/test.txt; the reference's `__getitem__` (mode "test": window start 550), `normalize` and
`denormalize` run as they are.  The tests rebuild the same inputs from the seed (`img_eagle_inputs`).
"""
import importlib.util
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SEED, T, H, W, WINDOW = 21, 600, 16, 32, 6


def img_eagle_inputs():
    rng = np.random.default_rng(SEED)
    states = (rng.standard_normal((T, H, W, 4)) * np.array([1.6, 1.9, 6.4, 9.1]) + np.array([0.0, 0.2, -0.5, 3.8])).astype(np.float32)
    pixel_type = rng.integers(0, 3, (H, W)).astype(np.int32)
    return states, pixel_type


def write_tree(tmp):
    states, pixel_type = img_eagle_inputs()
    d = os.path.join(tmp, "ds_img", "7", "1")
    os.makedirs(d)
    np.save(os.path.join(d, "states.npy"), states)
    np.save(os.path.join(d, "pixel_type.npy"), pixel_type)
    os.makedirs(os.path.join(tmp, "Splits"))
    with open(os.path.join(tmp, "Splits", "test.txt"), "w") as f:
        f.write("7/1\n")
    return os.path.join(tmp, "ds_img")


def main():
    spec = importlib.util.spec_from_file_location("ref_IMG_Eagle", "/root/reference/eagle/Dataloader/IMG_Eagle.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    tmp = tempfile.mkdtemp(prefix="fluidgrid_img_eagle_")
    data_path = write_tree(tmp)
    cwd = os.getcwd()
    os.chdir(tmp)                                   # the reference opens Splits/{mode}.txt relative to the working directory
    try:
        ds = mod.EagleDataset(data_path, mode="test", window_length=WINDOW)
        out = ds[0]
        den = ds.denormalize(out["states"])
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_img_eagle.npz"), states=out["states"].numpy(), mask=out["mask"],
                        example=out["example"].numpy(), denormalized=den.numpy(), n=np.array([len(ds)]))
    print("ref_img_eagle.npz", out["states"].shape, out["states"].dtype, out["example"])


if __name__ == "__main__":
    main()
