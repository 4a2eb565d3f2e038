"""SEVIR VIL sequences staged on the GPU: the step in front of the hot path.

``DeviceSEVIRLoader`` keeps the sampling contract of the reference's ``SEVIRDataLoader`` for the 'vil' type
(reference pipeline/datasets/sevir/sevir.py: shard bounds :346-375, ``__len__`` / ``use_up`` :534-560, zero padding
past the shard end :562-594, ``_sequent_sample`` :796-849, ``_idx_sample`` :851-908, ``preprocess_data_dict`` +
``change_layout`` :626-666, 88-101) but moves only RAW uint8 EVENTS over PCIe -- each event once, from pinned memory,
on a copy stream, one batch ahead of the consumer -- and cuts the ``seq_len`` windows, casts, rescales and lays them
out with one kernel (``wfk_stage_vil_windows_ex``). The reference slices on the host, converts to float32 there and
ships 4 bytes per pixel per window (windows of one event overlap: 3 x 25 frames out of 49).

``events`` is anything that is indexed like the HDF5 dataset the reference reads (``h5py.File(...)['vil']``,
sevir.py:562-566): ``events.shape == (E, H, W, raw_seq_len)``, ``events[i]`` -> uint8 [H, W, raw_seq_len]. A numpy
array, an ``np.memmap`` or a CPU uint8 tensor work as they are; reading the SEVIR catalogue and opening HDF5 files stay
with the caller (h5py / pandas are not part of this path).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import numpy.random as nprand
import torch

from . import _cabi

VALID_LAYOUT = ("NHWT", "NTHW", "NTCHW", "NTHWC", "TNHW", "TNCHW")      # sevir.py:226
VALID_SPLIT_MODE = ("ceil", "floor", "uneven")                          # sevir.py:232
RESCALE = {"01": (1 / 255, 0.0), "sevir": (1 / 47.54, -33.44)}          # 'vil' rows of sevir.py:44-63

Pick = Tuple[int, int]  # (event index into the shuffled sample list, sequence index inside the event)


class SamplePlan:
    """Host-side description of one batch: which (event, window) pairs, which events must be resident, the mask."""

    def __init__(self, picks: List[Pick], end_event_idx: int, stride: int):
        self.picks = picks
        self.events: List[int] = []            # distinct events in order of first use
        slot: Dict[int, int] = {}
        for e, _ in picks:
            if e not in slot:
                slot[e] = len(self.events)
                self.events.append(e)
        self.windows = [(slot[e], s * stride) for e, s in picks]      # (slot in the uploaded set, first raw frame)
        self.real = [e < end_event_idx for e in self.events]          # False: zero padding past the shard end
        mask = [e < end_event_idx for e, _ in picks]
        self.mask: Optional[List[bool]] = None if all(mask) else mask


class DeviceSEVIRLoader:
    def __init__(self, events, seq_len: int = 25, raw_seq_len: int = 49, sample_mode: str = "sequent", stride: int = 12,
                 batch_size: int = 1, layout: str = "NHWT", num_shard: int = 1, rank: int = 0, split_mode: str = "uneven",
                 shuffle: bool = False, shuffle_seed: int = 1, rescale_method: str = "01",
                 out_dtype: torch.dtype = torch.float32, device=None, prefetch: bool = True):
        if tuple(events.shape[3:]) != (raw_seq_len,) or len(events.shape) != 4:
            raise ValueError(f"events must be [E, H, W, raw_seq_len={raw_seq_len}], got {tuple(events.shape)}")
        assert seq_len <= raw_seq_len, f"seq_len must not be larger than raw_seq_len = {raw_seq_len}, got {seq_len}."
        assert sample_mode in ["random", "sequent"], f"Invalid sample_mode = {sample_mode}, must be 'random' or 'sequent'."
        if layout not in VALID_LAYOUT:
            raise ValueError(f"Invalid layout = {layout}! Must be one of {VALID_LAYOUT}.")
        if split_mode not in VALID_SPLIT_MODE:
            raise ValueError(f"Invalid split_mode: {split_mode}! Must be one of {VALID_SPLIT_MODE}.")
        if rescale_method not in RESCALE:
            raise ValueError(f"Invalid rescale option: {rescale_method}.")
        if out_dtype not in (torch.float32, torch.float16):
            raise ValueError("out_dtype must be float32 or float16")
        self.events = events
        self.seq_len, self.raw_seq_len, self.stride, self.batch_size = seq_len, raw_seq_len, stride, batch_size
        self.sample_mode, self.layout, self.num_shard, self.rank, self.split_mode = sample_mode, layout, num_shard, rank, split_mode
        self.shuffle, self.shuffle_seed = shuffle, int(shuffle_seed)
        self.rescale_method, self.out_dtype, self.prefetch = rescale_method, out_dtype, prefetch
        self.device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0) if device is None else torch.device(device)
        self._order = np.arange(int(events.shape[0]))
        if self.shuffle:
            self.shuffle_samples()
        self._pinned: List[Optional[torch.Tensor]] = [None, None]
        self._resident: List[Optional[torch.Tensor]] = [None, None]
        self._slot_free: List[Optional[torch.cuda.Event]] = [None, None]
        self._slot = 0
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._pending = None
        self.h2d_bytes = 0
        self.reset()  # like the reference constructor (sevir.py:271-273), this shuffles a second time when `shuffle`

    # ------------------------------------------------------------------ sampling contract (host logic, no GPU)
    def shuffle_samples(self):
        """``self._samples.sample(frac=1, random_state=seed)`` (sevir.py:307-308): pandas draws
        ``RandomState(seed).permutation(n)``; re-shuffling permutes the current order again, as in the reference."""
        perm = np.random.RandomState(self.shuffle_seed).permutation(len(self._order))
        self._order = self._order[perm]

    @property
    def total_num_event(self) -> int:
        return int(len(self._order))

    @property
    def start_event_idx(self) -> int:
        return self.total_num_event // self.num_shard * self.rank

    @property
    def end_event_idx(self) -> int:
        per = self.total_num_event // self.num_shard
        if self.split_mode == "ceil":
            return self.start_event_idx + (self.total_num_event - per * (self.num_shard - 1))
        if self.split_mode == "floor":
            return per * (self.rank + 1)
        return self.total_num_event if self.rank == self.num_shard - 1 else per * (self.rank + 1)

    @property
    def num_event(self) -> int:
        return self.end_event_idx - self.start_event_idx

    @property
    def num_seq_per_event(self) -> int:
        return 1 + (self.raw_seq_len - self.seq_len) // self.stride

    @property
    def total_num_seq(self) -> int:
        return int(self.num_seq_per_event * self.num_event)

    def __len__(self) -> int:
        return self.total_num_seq // self.batch_size

    @property
    def sample_count(self) -> int:
        return self._sample_count

    @property
    def curr_event_idx(self) -> int:
        return self._curr_event_idx

    @property
    def curr_seq_idx(self) -> int:
        return self._curr_seq_idx

    def reset(self, shuffle: Optional[bool] = None):
        self._curr_event_idx, self._curr_seq_idx, self._sample_count = self.start_event_idx, 0, 0
        self._pending = None
        if shuffle is None:
            shuffle = self.shuffle
        if shuffle:
            self.shuffle_samples()

    def _used_up_at(self, event_idx: int, seq_idx: int) -> bool:
        if self.sample_mode == "random":
            return False
        remain = (self.num_seq_per_event - seq_idx) + (self.end_event_idx - event_idx - 1) * self.num_seq_per_event
        return remain < self.batch_size if self.split_mode == "floor" else remain <= 0

    @property
    def use_up(self) -> bool:
        return self._used_up_at(self._curr_event_idx, self._curr_seq_idx)

    def _walk(self, event_idx: int, seq_idx: int) -> Tuple[List[Pick], int, int]:
        picks = []
        for _ in range(self.batch_size):
            picks.append((event_idx, seq_idx))
            seq_idx += 1
            if seq_idx >= self.num_seq_per_event:
                event_idx, seq_idx = event_idx + 1, 0
        return picks, event_idx, seq_idx

    def plan_sequent(self, event_idx: int, seq_idx: int) -> Tuple[SamplePlan, int, int]:
        picks, e, s = self._walk(event_idx, seq_idx)
        return SamplePlan(picks, self.end_event_idx, self.stride), e, s

    def plan_index(self, index: int) -> SamplePlan:
        """``_idx_sample`` addressing (sevir.py:864-877): batch ``index`` starts at sequence ``index * batch_size`` of the
        whole sample list (the reference ignores the shard start here; so does this)."""
        first = index * self.batch_size
        picks, _, _ = self._walk(first // self.num_seq_per_event, first % self.num_seq_per_event)
        return SamplePlan(picks, self.end_event_idx, self.stride)

    def plan_random(self) -> SamplePlan:
        """The draws of ``_random_sample`` (sevir.py:770-781: two ``numpy.random.randint`` calls on the global state).
        The reference's gather loop (:786-794) never advances ``num_sampled`` and does not terminate; this is the
        batch those draws describe."""
        ev = nprand.randint(low=self.start_event_idx, high=self.end_event_idx, size=self.batch_size)
        sq = nprand.randint(low=0, high=self.num_seq_per_event, size=self.batch_size)
        return SamplePlan([(int(e), int(s)) for e, s in zip(ev, sq)], self.end_event_idx, self.stride)

    # ------------------------------------------------------------------ device side
    def _begin_upload(self, plan: SamplePlan):
        """Gather the plan's events into pinned memory and start their copy on the copy stream."""
        n = len(plan.events)
        _, h, w, t = self.events.shape
        slot = self._slot
        self._slot ^= 1
        if self._pinned[slot] is None or self._pinned[slot].shape[0] < n:
            self._pinned[slot] = torch.empty((n, h, w, t), dtype=torch.uint8).pin_memory()
            self._resident[slot] = torch.empty((n, h, w, t), dtype=torch.uint8, device=self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        host, dev = self._pinned[slot], self._resident[slot]
        # the staging kernel that last read this slot's device buffer (two batches ago) must be done before the buffers
        # are reused; it ran after the slot's previous copy, so the pinned buffer is free as well
        if self._slot_free[slot] is not None:
            self._slot_free[slot].synchronize()
        host_np = host.numpy()
        for i, (e, real) in enumerate(zip(plan.events, plan.real)):
            if real:
                host_np[i] = np.asarray(self.events[int(self._order[e])], dtype=np.uint8)
            else:
                host_np[i] = 0
        with torch.cuda.stream(self._copy_stream):
            dev[:n].copy_(host[:n], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        self.h2d_bytes += n * h * w * t
        return dev, done, slot

    def _stage(self, plan: SamplePlan, dev: torch.Tensor, done: torch.cuda.Event, slot: int) -> torch.Tensor:
        lib = _cabi.init(self.device.index or 0)
        _, h, w, t_raw = self.events.shape
        n = len(plan.windows)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(done)
        win = torch.tensor(plan.windows, dtype=torch.int32).pin_memory().to(self.device, non_blocking=True)
        out = torch.empty((n, self.seq_len, 1, h, w), dtype=self.out_dtype, device=self.device)
        scale, offset = RESCALE[self.rescale_method]
        _cabi.check(lib.wfk_stage_vil_windows_ex(dev.data_ptr(), len(plan.events), h, w, t_raw, win.data_ptr(), n,
                                                 self.seq_len, scale, offset, out.data_ptr(),
                                                 0 if self.out_dtype == torch.float32 else 1, cur.cuda_stream),
                    "wfk_stage_vil_windows_ex")
        free = torch.cuda.Event()
        free.record(cur)
        self._slot_free[slot] = free
        return self._to_layout(out)

    def _to_layout(self, x: torch.Tensor) -> torch.Tensor:
        """[N, T, 1, H, W] -> ``self.layout`` (a view, like einops' rearrange in change_layout, sevir.py:88-101)."""
        if self.layout == "NTCHW":
            return x
        if self.layout == "TNCHW":
            return x.transpose(0, 1)
        y = x.squeeze(2)
        return {"NTHW": lambda: y, "NHWT": lambda: y.permute(0, 2, 3, 1), "NTHWC": lambda: y.unsqueeze(-1),
                "TNHW": lambda: y.transpose(0, 1)}[self.layout]()

    def _materialise(self, plan: SamplePlan) -> torch.Tensor:
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceSEVIRLoader stages on a B200 only (no CPU fallback)")
        return self._stage(plan, *self._begin_upload(plan))

    def _idx_sample(self, index: int) -> Dict[str, torch.Tensor]:
        return {"vil": self._materialise(self.plan_index(index))}

    def __getitem__(self, index: int) -> Dict[str, torch.Tensor]:
        return self._idx_sample(index)

    def __iter__(self):
        return self

    def __next__(self) -> Dict[str, object]:
        if self.sample_mode == "random":
            self._sample_count += 1
            return {"vil": self._materialise(self.plan_random())}
        if self.use_up:
            raise StopIteration
        self._sample_count += 1
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceSEVIRLoader stages on a B200 only (no CPU fallback)")
        if self._pending is None:
            plan, e, s = self.plan_sequent(self._curr_event_idx, self._curr_seq_idx)
            self._pending = (plan, e, s, *self._begin_upload(plan))
        plan, e, s, dev, done, slot = self._pending
        self._pending = None
        self._curr_event_idx, self._curr_seq_idx = e, s
        out = self._stage(plan, dev, done, slot)
        if self.prefetch and not self._used_up_at(e, s):
            nxt, e2, s2 = self.plan_sequent(e, s)
            self._pending = (nxt, e2, s2, *self._begin_upload(nxt))   # overlaps the caller's work on `out`
        return {"vil": out, "mask": plan.mask}


class DeviceSEVIRTorchDataset:
    """``SEVIRTorchDataset`` (sevir.py:981-1067) on top of ``DeviceSEVIRLoader``: batch size 1, ``__getitem__`` returns
    one sequence in ``layout`` (default "THWC") as a CUDA tensor. Augmentation modes other than "0" are torchvision
    transforms in the reference and are not part of this path."""

    def __init__(self, events, seq_len: int = 25, raw_seq_len: int = 49, sample_mode: str = "sequent", stride: int = 12,
                 layout: str = "THWC", split_mode: str = "uneven", shuffle: bool = False, shuffle_seed: int = 1,
                 rescale_method: str = "01", aug_mode: str = "0", ret_contiguous: bool = True, device=None,
                 out_dtype: torch.dtype = torch.float32):
        if aug_mode != "0":
            raise NotImplementedError("aug_mode other than '0' is out of scope (torchvision transforms)")
        self.layout = layout.replace("C", "1")
        self.ret_contiguous = ret_contiguous
        self.sevir_dataloader = DeviceSEVIRLoader(events, seq_len=seq_len, raw_seq_len=raw_seq_len, sample_mode=sample_mode,
                                                  stride=stride, batch_size=1, layout="NTCHW", num_shard=1, rank=0,
                                                  split_mode=split_mode, shuffle=shuffle, shuffle_seed=shuffle_seed,
                                                  rescale_method=rescale_method, device=device, out_dtype=out_dtype,
                                                  prefetch=False)

    def __getitem__(self, index: int) -> torch.Tensor:
        x = self.sevir_dataloader._idx_sample(index)["vil"].squeeze(0)  # [T, 1, H, W]
        src = {"T": 0, "1": 1, "H": 2, "W": 3}
        if sorted(self.layout) != sorted("T1HW"):
            raise ValueError(f"unsupported layout {self.layout}")
        x = x.permute(*[src[c] for c in self.layout])
        return x.contiguous() if self.ret_contiguous else x

    def __len__(self) -> int:
        return len(self.sevir_dataloader)
