"""Device-side input pipeline.

The reference's loaders parse one ASCII PLY with ``np.loadtxt`` and draw ``np.random.choice`` per sample on 4 host
workers (dataloader_multi_peak_vonMises.py:6-26,69-86; dataloader_8dir_sampled.py:6-15,41-58;
dataloader_single_peak_vonMises.py:5-14,47-52) - four to five orders of magnitude below what the training step consumes.
Here the dataset is parsed ONCE into a binary cache (``build_cache``), loaded into HBM as one ragged array
(``PointCloudCache``), and every batch is built by one CUDA launch that resamples the selected clouds to ``num_points``
(``pcoe_resample_clouds_f32``) - no per-step host work, no H2D copy of points, CUDA-graph capturable.

File parsers keep the reference's names and semantics: ``read_ply``, ``sample_pts`` (host version, for completeness),
``read_mvm_gt`` (= PointCloudDatasetMvM._read_mvM), ``read_vm_gt`` (= PointCloudDatasetVonMises._read_vm) and ``read_8dir_gt``
(the prob8 branch of dataloader_8dir_sampled.PointCloudDataset.__getitem__).

Cache file (little endian): magic ``PCOECACH`` | u32 version | u32 json_len | json header (cloud count, target kind, label
map, array table) | 64-byte aligned raw arrays (points f32 [total,3], offsets i64 [n+1], labels i64 [n], targets f32).
"""
from __future__ import annotations

import json
import os
import struct

import numpy as np
import torch

from . import _lib

_MAGIC = b"PCOECACH"
_VERSION = 1


# ------------------------------------------------------------------------------------------------
# parsers (reference semantics)
# ------------------------------------------------------------------------------------------------
def read_ply(p) -> np.ndarray:
    """ASCII PLY -> (n,3) float32: skip the header up to ``end_header``, then the first three columns of every row.
    Reference: dataloader_multi_peak_vonMises.py:6-19."""
    with open(p, "r") as f:
        while True:
            line = f.readline()
            if not line or line.strip() == "end_header":
                break
        pts = np.loadtxt(f, dtype=np.float32, ndmin=2)
    return np.ascontiguousarray(pts[:, :3]) if pts.size else np.zeros((0, 3), np.float32)


def sample_pts(arr: np.ndarray, num: int = 10000) -> np.ndarray:
    """Host version of the reference's resampling (dataloader_multi_peak_vonMises.py:21-26), numpy's generator."""
    n = len(arr)
    if n == 0:
        return arr
    return arr[np.random.choice(n, num, replace=(n < num))]


def read_mvm_gt(gt_path, max_K: int = 4):
    """-> ((max_K,3) float32 rows (mu, kappa, w) zero padded, K).  Reference: PointCloudDatasetMvM._read_mvM,
    dataloader_multi_peak_vonMises.py:35-67 (first non-comment line ``K <int>``, one header line, then data rows)."""
    with open(gt_path, "r", encoding="utf-8") as f:
        lines = [l.strip() for l in f if l.strip() and not l.startswith("#")]
    if len(lines) < 2:
        raise RuntimeError(f"GT file too short or malformed: {gt_path}")
    parts = lines[0].split()
    if len(parts) < 2:
        raise RuntimeError(f"GT file K line malformed: {gt_path}")
    K = int(parts[1])
    rows = []
    for ln in lines[2:]:
        vals = ln.split()
        if len(vals) >= 3:
            rows.append([float(vals[0]), float(vals[1]), float(vals[2])])
    while len(rows) < max_K:
        rows.append([0.0, 0.0, 0.0])
    return np.asarray(rows, dtype=np.float32)[:max_K], K


def read_vm_gt(path):
    """-> (mu, kappa >= 0); (0, 0) when the file is missing or malformed.  Reference: PointCloudDatasetVonMises._read_vm,
    dataloader_single_peak_vonMises.py:36-45 (first non-comment line)."""
    try:
        with open(path, "r", encoding="utf-8") as f:
            lines = [l.strip() for l in f if l.strip() and not l.startswith("#")]
        mu, kappa = map(float, lines[0].split()[:2])
    except Exception:
        mu, kappa = 0.0, 0.0
    return mu, max(kappa, 0.0)


def read_8dir_gt(prob_path, uniform: bool = False) -> np.ndarray:
    """-> (8,) float32; uniform 0.125 when the class is in the uniform set, the file is missing or unreadable.
    Reference: dataloader_8dir_sampled.py:48-56."""
    if uniform or not os.path.exists(prob_path):
        return np.full(8, 0.125, dtype=np.float32)
    try:
        return np.loadtxt(prob_path, dtype=np.float32).flatten()[:8].astype(np.float32)
    except Exception:
        return np.full(8, 0.125, dtype=np.float32)


# ------------------------------------------------------------------------------------------------
# binary cache
# ------------------------------------------------------------------------------------------------
def build_cache(samples, out_path: str, kind: str = "mvm", max_K: int = 4, label_map=None, uniform_set=()) -> dict:
    """Parse every sample once and write the binary cache.

    kind 'mvm'  : samples = [(ply_path, gt_txt_path, category)]  -> targets (n,max_K,3) + K (n,)      (PointCloudDatasetMvM)
    kind 'vm'   : samples = [(ply_path, category)], GT next to the PLY as ``<stem>_single_peak_vM_gt.txt`` -> (n,2)
    kind '8dir' : samples = [(ply_path, prob8_path, category)]   -> (n,8)
    The label map follows the reference (sorted categories for 'mvm', first-seen order otherwise)."""
    samples = list(samples)
    cats = [s[-1] for s in samples]
    if label_map is None:
        if kind == "mvm":
            label_map = {c: i for i, c in enumerate(sorted(set(cats)))}
        else:
            label_map = {}
            for c in cats:
                label_map.setdefault(c, len(label_map))
    pts, offsets, targets, Ks = [], [0], [], []
    for s in samples:
        p = read_ply(s[0])
        pts.append(p)
        offsets.append(offsets[-1] + len(p))
        if kind == "mvm":
            t, K = read_mvm_gt(s[1], max_K)
            targets.append(t); Ks.append(K)
        elif kind == "vm":
            stem = os.path.splitext(str(s[0]))[0]
            targets.append(np.asarray(read_vm_gt(stem + "_single_peak_vM_gt.txt"), dtype=np.float32))
        elif kind == "8dir":
            targets.append(read_8dir_gt(s[1], s[2] in set(uniform_set)))
        else:
            raise ValueError(f"unknown cache kind {kind!r}")
    arrays = {
        "points": np.concatenate(pts, 0).astype(np.float32) if pts else np.zeros((0, 3), np.float32),
        "offsets": np.asarray(offsets, dtype=np.int64),
        "labels": np.asarray([label_map[c] for c in cats], dtype=np.int64),
        "targets": np.stack(targets).astype(np.float32) if targets else np.zeros((0,), np.float32),
    }
    if kind == "mvm":
        arrays["K"] = np.asarray(Ks, dtype=np.int64)
    table, blobs, cur = {}, [], 0
    for name, a in arrays.items():
        a = np.ascontiguousarray(a)
        cur = (cur + 63) // 64 * 64
        table[name] = {"dtype": str(a.dtype), "shape": list(a.shape), "offset": cur}
        blobs.append((cur, a.tobytes()))
        cur += a.nbytes
    header = json.dumps({"n": len(samples), "kind": kind, "max_K": max_K, "label_map": label_map, "arrays": table}).encode()
    with open(out_path, "wb") as f:
        f.write(_MAGIC + struct.pack("<II", _VERSION, len(header)) + header)
        base = (f.tell() + 63) // 64 * 64
        for off, blob in blobs:
            f.seek(base + off)
            f.write(blob)
    return {"n": len(samples), "points": int(offsets[-1]), "bytes": base + cur}


class PointCloudCache:
    """A cached dataset resident on one GPU.  ``batch(cloud_ids, num_points)`` is the device replacement of
    ``default_collate([dataset[i] for i in ids])``."""

    def __init__(self, arrays: dict, meta: dict, device):
        self.kind, self.max_K, self.label_map = meta["kind"], meta["max_K"], meta["label_map"]
        self.n = meta["n"]
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pcoe: PointCloudCache lives on a CUDA device (no CPU path)")
        up = lambda a: torch.from_numpy(a).pin_memory().to(self.device, non_blocking=True)
        self.points = up(arrays["points"].reshape(-1, 3))
        self.offsets = up(arrays["offsets"])
        self.labels = up(arrays["labels"])
        self.targets = up(arrays["targets"])
        self.K = up(arrays["K"]) if "K" in arrays else None
        self.sizes = np.diff(arrays["offsets"])
        self._draws = 0

    @classmethod
    def load(cls, path: str, device="cuda"):
        with open(path, "rb") as f:
            if f.read(8) != _MAGIC:
                raise ValueError(f"{path}: not a pcoe point-cloud cache")
            version, hlen = struct.unpack("<II", f.read(8))
            if version != _VERSION:
                raise ValueError(f"{path}: cache version {version}, expected {_VERSION}")
            meta = json.loads(f.read(hlen))
            base = (f.tell() + 63) // 64 * 64
            arrays = {}
            for name, t in meta["arrays"].items():
                f.seek(base + t["offset"])
                cnt = int(np.prod(t["shape"])) if t["shape"] else 1
                arrays[name] = np.fromfile(f, dtype=np.dtype(t["dtype"]), count=cnt).reshape(t["shape"])
        return cls(arrays, meta, device)

    @classmethod
    def from_arrays(cls, clouds, targets, labels=None, kind="mvm", K=None, max_K=4, device="cuda"):
        """In-memory construction (synthetic data, tests): ``clouds`` = list of (n_i,3) arrays."""
        offsets = np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).astype(np.int64)
        arrays = {"points": np.concatenate(clouds, 0).astype(np.float32), "offsets": offsets,
                  "labels": np.asarray(labels if labels is not None else np.zeros(len(clouds)), dtype=np.int64),
                  "targets": np.asarray(targets, dtype=np.float32)}
        if K is not None:
            arrays["K"] = np.asarray(K, dtype=np.int64)
        return cls(arrays, {"kind": kind, "max_K": max_K, "label_map": {}, "n": len(clouds)}, device)

    def __len__(self):
        return self.n

    def batch(self, cloud_ids: torch.Tensor, num_points: int, seed: int = 0, counter: torch.Tensor | None = None,
              return_idx: bool = False, draw: int | None = None):
        """cloud_ids (B,) int -> (xyz (B,num_points,3), targets[, K], labels[, src_idx]) on the device.
        ``draw``: explicit draw number of element 0 (default: a running counter, so repeated calls differ);
        ``counter``: 1-element int64 CUDA tensor added (times B) on the device - for CUDA-graph replays."""
        ids = cloud_ids.to(self.device, torch.int32).contiguous()
        B = ids.numel()
        xyz = torch.empty(B, num_points, 3, dtype=torch.float32, device=self.device)
        idx = torch.empty(B, num_points, dtype=torch.int32, device=self.device) if return_idx else None
        base = self._draws if draw is None else draw
        if draw is None:
            self._draws += B
        _lib.check(_lib.load().pcoe_resample_clouds_f32(
            self.points.data_ptr(), self.offsets.data_ptr(), self.n, ids.data_ptr(), B, num_points, seed & (2**64 - 1),
            base, None if counter is None else counter.data_ptr(), xyz.data_ptr(), None if idx is None else idx.data_ptr(),
            torch.cuda.current_stream().cuda_stream))
        il = ids.long()
        out = [xyz, self.targets.index_select(0, il)]
        if self.K is not None:
            out.append(self.K.index_select(0, il))
        out.append(self.labels.index_select(0, il))
        if return_idx:
            out.append(idx)
        return tuple(out)


class DeviceLoader:
    """Iterates a ``PointCloudCache`` like ``DataLoader(dataset, batch_size, shuffle, drop_last)``: one epoch = one pass
    over a (shuffled) permutation of the clouds; every batch is built on the device."""

    def __init__(self, cache: PointCloudCache, batch_size: int, num_points: int, shuffle: bool = True,
                 drop_last: bool = False, seed: int = 0):
        self.cache, self.batch_size, self.num_points = cache, batch_size, num_points
        self.shuffle, self.drop_last, self.seed = shuffle, drop_last, seed
        self.epoch = 0

    def __len__(self):
        n = len(self.cache)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.cache)
        if self.shuffle:
            g = torch.Generator(device=self.cache.device).manual_seed(self.seed + self.epoch)
            order = torch.randperm(n, device=self.cache.device, generator=g)
        else:
            order = torch.arange(n, device=self.cache.device)
        self.epoch += 1
        for i in range(len(self)):
            ids = order[i * self.batch_size:(i + 1) * self.batch_size]
            yield self.cache.batch(ids, self.num_points, seed=self.seed)
