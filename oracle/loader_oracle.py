"""TEST INFRASTRUCTURE ONLY (imported by tests/ and the golden generator, never by the product path).

numpy restatement of the 'vil' sampling of ``SEVIRDataLoader`` (reference pipeline/datasets/sevir/sevir.py) over
an in-memory uint8 event array [E, H, W, raw_seq_len] standing in for the HDF5 dataset ``f['vil']``:
shard bounds (:346-375), ``__len__`` / ``use_up`` (:534-560), ``_load_event_batch`` zero padding (:562-594),
``_sequent_sample`` (:796-849), ``_idx_sample`` (:851-908), ``_random_sample`` (:770-794) and
``preprocess_data_dict`` + ``change_layout`` (:626-666, 88-101).

Pinned: tests/golden/loader_golden.npz holds outputs of the UNMODIFIED reference class (HDF5 handle replaced by
the same in-memory array; tests/golden/make_golden_loader.py) and tests/test_loader.py checks this file against
them bit for bit.
"""
import numpy as np
import numpy.random as nprand

SCALE = {"01": (1 / 255, 0), "sevir": (1 / 47.54, -33.44)}  # sevir.py:44-63 ('vil' rows)
_EINOPS = {"NHWT": (0, 1, 2, 3), "NTHW": (0, 3, 1, 2), "TNHW": (3, 0, 1, 2)}


def change_layout(x_nhwt: np.ndarray, layout: str) -> np.ndarray:
    """einops ``rearrange('N H W T -> <layout with C as 1>')`` (sevir.py:88-101)."""
    if layout in _EINOPS:
        return x_nhwt.transpose(_EINOPS[layout])
    if layout == "NTCHW":
        return x_nhwt.transpose(0, 3, 1, 2)[:, :, None]
    if layout == "NTHWC":
        return x_nhwt.transpose(0, 3, 1, 2)[..., None]
    if layout == "TNCHW":
        return x_nhwt.transpose(3, 0, 1, 2)[:, :, None]
    raise ValueError(layout)


def preprocess(x_nhwt: np.ndarray, layout: str = "NHWT", rescale: str = "01") -> np.ndarray:
    """``scale * (x.float() + offset)``: python-float scalars applied to a float32 tensor, i.e. float32 arithmetic."""
    scale, offset = SCALE[rescale]
    x = x_nhwt.astype(np.float32)
    return change_layout(np.float32(scale) * (x + np.float32(offset)), layout)


class LoaderOracle:
    def __init__(self, events, seq_len=25, raw_seq_len=49, sample_mode="sequent", stride=12, batch_size=1,
                 layout="NHWT", num_shard=1, rank=0, split_mode="uneven", rescale="01", order=None):
        self.events = events
        self.order = list(range(events.shape[0])) if order is None else list(order)
        self.seq_len, self.raw_seq_len, self.stride, self.batch_size = seq_len, raw_seq_len, stride, batch_size
        self.sample_mode, self.layout, self.num_shard, self.rank = sample_mode, layout, num_shard, rank
        self.split_mode, self.rescale = split_mode, rescale
        self.reset()

    # ---- shard bounds (sevir.py:346-375)
    @property
    def total_num_event(self):
        return len(self.order)

    @property
    def start_event_idx(self):
        return self.total_num_event // self.num_shard * self.rank

    @property
    def end_event_idx(self):
        per = self.total_num_event // self.num_shard
        if self.split_mode == "ceil":
            return self.start_event_idx + self.total_num_event - per * (self.num_shard - 1)
        if self.split_mode == "floor" or self.rank != self.num_shard - 1:
            return per * (self.rank + 1)
        return self.total_num_event

    @property
    def num_seq_per_event(self):
        return 1 + (self.raw_seq_len - self.seq_len) // self.stride

    def __len__(self):
        return int(self.num_seq_per_event * (self.end_event_idx - self.start_event_idx)) // self.batch_size

    def reset(self):
        self.curr_event_idx, self.curr_seq_idx = self.start_event_idx, 0

    @property
    def use_up(self):
        if self.sample_mode == "random":
            return False
        remain = (self.num_seq_per_event - self.curr_seq_idx) + \
            (self.end_event_idx - self.curr_event_idx - 1) * self.num_seq_per_event
        return remain < self.batch_size if self.split_mode == "floor" else remain <= 0

    # ---- event loading with zero padding past the shard end (sevir.py:562-594)
    def _load(self, event_idx, count):
        stop = min(event_idx + count, self.end_event_idx)
        got = [self.events[self.order[i]] for i in range(event_idx, stop)]
        h, w, t = self.events.shape[1:]
        pad = [np.zeros((h, w, t), dtype=self.events.dtype)] * (event_idx + count - stop)
        return np.stack(got + pad, axis=0)

    def _walk(self, event_idx, seq_idx):
        picks = []
        for _ in range(self.batch_size):
            picks.append((event_idx, seq_idx))
            seq_idx += 1
            if seq_idx >= self.num_seq_per_event:
                event_idx, seq_idx = event_idx + 1, 0
        return picks, event_idx, seq_idx

    def _gather(self, picks):
        start = picks[0][0]
        batch = self._load(start, picks[-1][0] - start + 1)
        seqs = [batch[e - start, :, :, s * self.stride:s * self.stride + self.seq_len] for e, s in picks]
        return preprocess(np.stack(seqs, axis=0), self.layout, self.rescale)

    def sequent_sample(self):
        assert not self.use_up
        picks, e, s = self._walk(self.curr_event_idx, self.curr_seq_idx)
        mask = [ev < self.end_event_idx for ev, _ in picks]
        self.curr_event_idx, self.curr_seq_idx = e, s
        return self._gather(picks), (None if all(mask) else mask)

    def idx_sample(self, index):
        e = (index * self.batch_size) // self.num_seq_per_event
        s = (index * self.batch_size) % self.num_seq_per_event
        picks, _, _ = self._walk(e, s)
        return self._gather(picks)

    def random_sample(self):
        ev = nprand.randint(low=self.start_event_idx, high=self.end_event_idx, size=self.batch_size)
        sq = nprand.randint(low=0, high=self.num_seq_per_event, size=self.batch_size)
        seqs = [self._load(int(e), 1)[0, :, :, int(s) * self.stride:int(s) * self.stride + self.seq_len]
                for e, s in zip(ev, sq)]
        return preprocess(np.stack(seqs, axis=0), self.layout, self.rescale)

    def __iter__(self):
        return self

    def __next__(self):
        if self.sample_mode == "random":
            return self.random_sample(), None
        if self.use_up:
            raise StopIteration
        return self.sequent_sample()
