#!/usr/bin/env python3
"""Golden vectors for the loader-side neighbour gather: runs the UNMODIFIED
reference dataset class my_dataloader/modelnet_with_lrf_sample_index_loader.py::ModelNetDataset on a
tiny synthetic dataset written to a temp directory in the file formats it expects (.txt points+normals,
.qua LRF quaternions, .ds{N}.pt pooling indices, .idx invalid-LRF ids), and stores the gather's inputs
and outputs.  Build container only (needs the reference checkout):

    python tests/golden/make_golden_gather.py [/path/to/reference]

`pyquaternion` (imported but unused by the class) is absent here and is stubbed."""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def main():
    stub = types.ModuleType("pyquaternion")
    stub.Quaternion = object
    sys.modules["pyquaternion"] = stub
    sys.path.insert(0, os.path.join(REF, "models"))
    sys.path.insert(0, os.path.join(REF, "my_dataloader"))
    import modelnet_with_lrf_sample_index_loader as ref_loader

    rng = np.random.default_rng(5)
    out = {}
    with tempfile.TemporaryDirectory() as root:
        names = ["chair", "lamp"]
        open(os.path.join(root, "modelnet10_shape_names.txt"), "w").write("\n".join(names) + "\n")
        ids = ["chair_0001", "lamp_0001", "lamp_0002"]
        open(os.path.join(root, "modelnet10_train.txt"), "w").write("\n".join(ids) + "\n")
        open(os.path.join(root, "modelnet10_test.txt"), "w").write(ids[0] + "\n")
        num_gen = 3
        N = 1500
        for sid in ids:
            d = os.path.join(root, "_".join(sid.split("_")[:-1]))
            os.makedirs(d, exist_ok=True)
            pts = rng.standard_normal((N, 6)).astype(np.float32)
            np.savetxt(os.path.join(d, sid + ".txt"), pts, delimiter=",", fmt="%.6f")
            q = rng.standard_normal((N, 4)).astype(np.float32)
            q /= np.linalg.norm(q, axis=1, keepdims=True)
            np.savetxt(os.path.join(d, sid + ".qua"), q, fmt="%.6f")
            ds = torch.full((num_gen, 1024 + 256 + 1, 9), -1, dtype=torch.int64)
            for g in range(num_gen):
                n1 = 900 + 40 * g                                  # pool1 centres in use
                p1 = torch.from_numpy(rng.integers(0, N, (n1, 9)))
                p1[torch.from_numpy(rng.random((n1, 9)) < 0.15)] = -1
                p1[:, 0] = torch.from_numpy(rng.integers(1, N, n1))  # a centre always exists
                ds[g, :n1] = p1
                n2 = 200 + 10 * g                                  # pool2 centres in use
                p2 = torch.from_numpy(rng.integers(0, n1, (n2, 9)))
                p2[torch.from_numpy(rng.random((n2, 9)) < 0.2)] = -1
                p2[:, 0] = torch.from_numpy(rng.integers(1, n1, n2))
                ds[g, 1024:1024 + n2] = p2
                ds[g, 1280, 0] = n1
            torch.save(ds, os.path.join(d, sid + ".ds%d.pt" % num_gen))
            wrong = rng.integers(0, N, 7) if sid != "lamp_0002" else np.zeros((0,), dtype=np.int64)
            np.savetxt(os.path.join(d, sid + ".idx"), wrong, fmt="%d")

        dataset = ref_loader.ModelNetDataset(root=root, npoints=64, split="train", num_of_class=10,
                                             num_gen_samples=num_gen, data_aug=False)
        for n, sid in enumerate(ids):
            np.random.seed(100 + n)  # fixes the reference's np.random.choice of the index set
            d = os.path.join(root, "_".join(sid.split("_")[:-1]))
            raw_pts = np.loadtxt(os.path.join(d, sid + ".txt"), delimiter=",").astype(np.float32)[:, :3]
            raw_lrf = np.loadtxt(os.path.join(d, sid + ".qua")).astype(np.float32)
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                wrong = np.loadtxt(os.path.join(d, sid + ".idx"), ndmin=1).astype(np.int64)
            points, lrfs, act, idx, _, cls = dataset[n]
            out["s%d.point_set" % n] = raw_pts
            out["s%d.lrf_set" % n] = raw_lrf
            out["s%d.wrong_ids" % n] = wrong
            out["s%d.idx" % n] = idx.numpy()
            out["s%d.points" % n] = points.numpy()
            out["s%d.lrfs" % n] = lrfs.numpy()
            out["s%d.act" % n] = act.numpy()
            print(sid, "idx", tuple(idx.shape), "active", int(act.sum()), "empties", int((idx < 0).sum()))
    np.savez_compressed(os.path.join(HERE, "gather.npz"), **out)


if __name__ == "__main__":
    main()
