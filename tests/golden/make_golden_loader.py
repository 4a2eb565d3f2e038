"""Golden fixtures for the SEVIR sampling contract, generated from the UNMODIFIED reference class
``SEVIRDataLoader`` (/root/reference/pipeline/datasets/sevir/sevir.py; build container only). Run:

    python tests/golden/make_golden_loader.py           # writes loader_golden.npz

The reference module imports h5py, lightning and matplotlib at the top (absent here, none used by the sampling
code): empty stand-in modules satisfy the imports. The one patch is ``_open_files``: the HDF5 handle
``h5py.File(...)`` is replaced by a dict holding the same in-memory uint8 array the tests regenerate from the seed,
which supports the only access the loader makes (``f['vil'][idx:idx+1, :, :, slice]``, sevir.py:562-566).
"""
import os
import sys
import types

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

for _name, _attrs in (("h5py", {"File": None}),
                      ("lightning", {"LightningDataModule": object, "seed_everything": lambda *a, **k: None}),
                      ("matplotlib", {}), ("matplotlib.colors", {"ListedColormap": None, "BoundaryNorm": None})):
    _m = types.ModuleType(_name)
    for _k, _v in _attrs.items():
        setattr(_m, _k, _v)
    sys.modules[_name] = _m

from pipeline.datasets.sevir.sevir import SEVIRDataLoader, SEVIRTorchDataset  # noqa: E402

from weatherforecastingtoolkit_b200.synthetic import make_loader_events  # noqa: E402

E, H, W, T_RAW = 5, 8, 12, 49
# every case: loader kwargs; `iterate` records the whole sequential pass, `index` one _idx_sample call
CASES = {
    "b4_uneven_ntchw": dict(batch_size=4, layout="NTCHW", split_mode="uneven"),
    "b4_floor_nhwt": dict(batch_size=4, layout="NHWT", split_mode="floor"),
    "b2_ceil_nthwc_shard1of2": dict(batch_size=2, layout="NTHWC", split_mode="ceil", num_shard=2, rank=1),
    "b3_uneven_tnhw_shard0of2": dict(batch_size=3, layout="TNHW", split_mode="uneven", num_shard=2, rank=0),
    "b4_sevir_rescale": dict(batch_size=4, layout="NTHW", split_mode="uneven", rescale_method="sevir"),
    "b5_stride6_len13": dict(batch_size=5, layout="TNCHW", split_mode="uneven", seq_len=13, stride=6),
    "b4_shuffled": dict(batch_size=4, layout="NTCHW", split_mode="uneven", shuffle=True, shuffle_seed=3),
}


def catalog():
    return pd.DataFrame({"id": [f"ev{i}" for i in range(E)], "img_type": ["vil"] * E, "file_name": ["f0.h5"] * E,
                         "file_index": list(range(E)), "time_utc": pd.to_datetime(["2019-01-01"] * E),
                         "pct_missing": [0] * E})


def main():
    events = make_loader_events(E, H, W, T_RAW, seed=7).numpy()

    def _open_files(self, verbose=True):  # the IO stub: stands in for h5py.File(...) (sevir.py:318-330)
        self._hdf_files = {"f0.h5": {"vil": events}}

    SEVIRDataLoader._open_files = _open_files
    InMemory = SEVIRDataLoader

    out = {}
    for name, kw in CASES.items():
        kw = dict(kw)
        args = dict(data_types=["vil"], seq_len=kw.pop("seq_len", 25), raw_seq_len=T_RAW, sample_mode="sequent",
                    stride=kw.pop("stride", 12), sevir_catalog=catalog(), sevir_data_dir="/nonexistent", verbose=False)
        args.update(kw)
        ld = InMemory(**args)
        out[f"{name}/len"] = np.int64(len(ld))
        out[f"{name}/order"] = np.asarray([int(np.ravel(i)[0]) for i in ld._samples["vil_index"].values], dtype=np.int64)
        nb = 0
        for d in ld:
            out[f"{name}/batch{nb}"] = d["vil"].numpy().copy()
            m = d["mask"]
            out[f"{name}/mask{nb}"] = np.ones(args["batch_size"], dtype=bool) if m is None else np.asarray(m, dtype=bool)
            out[f"{name}/mask{nb}_is_none"] = np.bool_(m is None)
            nb += 1
        out[f"{name}/num_batches"] = np.int64(nb)
        out[f"{name}/idx1"] = ld._idx_sample(1)["vil"].numpy().copy()

    # SEVIRTorchDataset.__getitem__ (sevir.py:1052-1064), default layout "THWC"
    ds = SEVIRTorchDataset(seq_len=25, raw_seq_len=T_RAW, stride=12, layout="THWC", sevir_catalog=catalog(),
                           sevir_data_dir="/nonexistent", verbose=False)
    out["dataset/len"] = np.int64(len(ds))
    for i in (0, 4, len(ds) - 1):
        out[f"dataset/item{i}"] = ds[i].numpy().copy()
    path = os.path.join(HERE, "loader_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
