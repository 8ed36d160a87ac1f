"""Input pipeline: trials -> pinned host memory -> device, prefetched ahead of the training loop.

Reference: ``behavenet/data/data_generator.py`` (``split_trials`` 42-103, ``SingleSessionDatasetBatchedLoad``
137-343, ``ConcatSessionsGenerator`` 432-633).  The reference serves one trial per ``next_batch`` call from a
``DataLoader(batch_size=1, num_workers=0, pin_memory=False)``: the HDF5 file is opened, the trial is read,
converted to float32 and divided by 255 on the host, and only then copied to the GPU, all on the training
thread.  At B200 step rates that loop is the bottleneck, so here

* the order of an epoch is drawn up front (same process as the reference: a session is chosen with
  probability ``batch_ratios``, exhausted sessions are skipped, trials inside a session come in a uniformly
  random order), which lets a worker thread run ``depth`` trials ahead of the consumer;
* the worker gathers each trial into a reusable pinned buffer and issues the host->device copy on a side
  stream; ``next_batch`` only makes the compute stream wait on that copy's event;
* video stays uint8 across PCIe (a quarter of the bytes) and is scaled by 1/255 on the device -- or not at
  all with ``raw_uint8=True``, for the encode-only path whose first layer reads bytes (``bn_cae_encode_u8``);
* with ``shard_frames=True`` every rank of ``behavenet_b200.parallel`` stages only its own contiguous frame
  range of each trial and the batch carries ``data['shard'] = (first frame, frames in the trial)``, the
  form ``AE.loss`` accepts.

The object keeps the reference generator's protocol (``datasets[i].{n_trials, batch_idxs, n_batches, lab,
expt, animal, session}``, ``n_datasets``, ``batch_ratios``, ``n_tot_batches``, ``reset_iterators``,
``next_batch`` returning ``(sample, dataset)`` with a leading batch axis of 1), so the reference's ``fit()``
loop (``fitting/training.py:316-390``) and ``export_latents`` run on it unchanged.
"""

import os
import queue
import threading
from collections import OrderedDict

import numpy as np

_SPLITS = ('train', 'val', 'test')
_FLOAT_SIGNALS = ('images', 'masks', 'neural', 'labels', 'labels_sc', 'labels_masks', 'ae_latents', 'latents',
                  'ae_predictions', 'arhmm_predictions')
_INT_SIGNALS = ('arhmm', 'arhmm_states')


def split_trials(n_trials, rng_seed=0, train_tr=8, val_tr=1, test_tr=1, gap_tr=0):
    """Blocked train/val/test split, ``train | gap | val | gap | test | gap`` per block, blocks shuffled.

    Draws from numpy's global generator in the reference's order (seed, offset, block permutation;
    data_generator.py:72-101), so a given ``rng_seed`` yields the reference's split."""
    block = train_tr + val_tr + test_tr + 3 * gap_tr
    n_blocks = n_trials // block
    if n_blocks == 0:
        raise ValueError('Not enough trials (n=%i) for the train/test/val/gap values %i/%i/%i/%i' %
                         (n_trials, train_tr, val_tr, test_tr, gap_tr))
    np.random.seed(rng_seed)
    spare = n_trials - n_blocks * block
    offset = np.random.randint(0, high=spare) if spare > 0 else 0
    order = np.random.permutation(n_blocks)
    starts = order * block + offset
    val0 = train_tr + gap_tr
    test0 = val0 + val_tr + gap_tr
    within = {'train': np.arange(train_tr), 'val': val0 + np.arange(val_tr), 'test': test0 + np.arange(test_tr)}
    return {k: (starts[:, None] + w[None, :]).reshape(-1) for k, w in within.items()}


class ArraySource:
    """Trials held in memory: ``signals`` maps a signal name to a list of per-trial arrays."""

    def __init__(self, signals, lab='', expt='', animal='', session=''):
        self.signals = list(signals.keys())
        self._data = signals
        counts = {len(v) for v in signals.values()}
        if len(counts) != 1:
            raise ValueError('every signal needs the same number of trials, got %s' % sorted(counts))
        self.n_trials = counts.pop()
        self.lab, self.expt, self.animal, self.session = lab, expt, animal, session

    def trial_length(self, idx):
        return int(self._data[self.signals[0]][idx].shape[0])

    def load(self, signal, idx, lo=0, hi=None):
        # a fresh copy per call, like the reference's per-trial HDF5 read: in-place transforms (ZScore,
        # Threshold, ...) run on the returned array in the worker thread and must not touch the caller's data
        return np.array(self._data[signal][idx][lo:hi], copy=True)


class HDF5Source:
    """One session's ``data.hdf5`` in the reference layout: group ``<signal>/trial_%04i`` per trial
    (data_generator.py:253-303).  The file is opened once per worker thread, not once per trial.

    Read through ``h5py`` when it is installed (like the reference) and through the dependency-free reader
    ``behavenet_b200.data.hdf5_lite`` otherwise (``backend='lite'`` forces the latter); both expose the same
    calls."""

    def __init__(self, path, signals, lab='', expt='', animal='', session='', backend=None):
        if backend not in (None, 'h5py', 'lite'):
            raise ValueError('backend must be None, "h5py" or "lite"')
        h5 = None
        if backend != 'lite':
            try:
                import h5py as h5
            except ImportError:
                if backend == 'h5py':
                    raise
        if h5 is None:
            from behavenet_b200.data import hdf5_lite as h5
        self._h5 = h5
        self.backend = 'lite' if h5.__name__.endswith('hdf5_lite') else 'h5py'
        self.path = path
        self.signals = list(signals)
        self.lab, self.expt, self.animal, self.session = lab, expt, animal, session
        self._local = threading.local()
        with h5.File(path, 'r', libver='latest', swmr=True) as f:
            for s in self.signals:
                if s not in f:
                    raise KeyError('%s has no group "%s"' % (path, s))
            self.n_trials = len(f[self.signals[0]])

    def __getstate__(self):
        st = dict(self.__dict__)
        st.pop('_local')
        st['_h5'] = self._h5.__name__
        return st

    def __setstate__(self, st):
        import importlib
        self.__dict__.update(st)
        self._h5 = importlib.import_module(st['_h5'])
        self._local = threading.local()

    def _file(self):
        f = getattr(self._local, 'f', None)
        if f is None:
            f = self._local.f = self._h5.File(self.path, 'r', libver='latest', swmr=True)
        return f

    def trial_length(self, idx):
        return int(self._file()[self.signals[0]]['trial_%04i' % idx].shape[0])

    def load(self, signal, idx, lo=0, hi=None):
        return self._file()[signal]['trial_%04i' % idx][lo:hi]


class _Session:
    """What the reference keeps per ``SingleSessionDatasetBatchedLoad``."""

    def __init__(self, source):
        self.source = source
        self.signals = list(source.signals)
        self.n_trials = source.n_trials
        self.lab, self.expt, self.animal, self.session = source.lab, source.expt, source.animal, source.session
        self.name = os.path.join(self.lab, self.expt, self.animal, self.session)
        self.sess_str = '%s_%s_%s_%s' % (self.lab, self.expt, self.animal, self.session)
        self.batch_idxs = None
        self.n_batches = None

    def __len__(self):
        return self.n_trials


class _PinnedSlot:
    """Grow-only pinned buffers of one in-flight trial, guarded by the event of its last copy."""

    def __init__(self):
        self.bufs = {}
        self.event = None

    def buffer(self, signal, nbytes):
        import torch
        b = self.bufs.get(signal)
        if b is None or b.numel() < nbytes:
            b = self.bufs[signal] = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8).pin_memory()
        return b[:nbytes]


class PrefetchSessionsGenerator:
    """Drop-in for the reference's ``ConcatSessionsGenerator`` over ``ArraySource`` / ``HDF5Source`` sessions."""

    _dtypes = {'train', 'val', 'test'}

    def __init__(self, sources, device='cuda', as_numpy=False, rng_seed=0, trial_splits=None, train_frac=1.0,
                 depth=3, raw_uint8=False, shard_frames=False, transforms=None):
        self.datasets = [_Session(s) for s in sources]
        self.n_datasets = len(self.datasets)
        self.datasets_info = [{'lab': d.lab, 'expt': d.expt, 'animal': d.animal, 'session': d.session}
                              for d in self.datasets]
        self.device = device
        self.as_numpy = as_numpy
        self.batch_load = True
        self.depth = max(1, int(depth))
        self.raw_uint8 = raw_uint8
        self.shard_frames = shard_frames
        self.transforms = transforms or {}
        if trial_splits is None:
            trial_splits = {'train_tr': 8, 'val_tr': 1, 'test_tr': 1, 'gap_tr': 0}
        ratios = []
        for ds in self.datasets:
            ds.batch_idxs = split_trials(len(ds), rng_seed=rng_seed, **trial_splits)
            if train_frac != 1.0:                                  # data_generator.py:520-537
                n_b = len(ds.batch_idxs['train'])
                if train_frac < 1.0:
                    n_keep = int(np.floor(train_frac * n_b))
                    if n_keep <= 0:
                        print('warning: attempting to use invalid number of training batches; '
                              'defaulting to all training batches')
                        n_keep = n_b
                else:
                    n_keep = int(min(train_frac, n_b))
                ds.batch_idxs['train'] = ds.batch_idxs['train'][np.random.choice(n_b, size=n_keep, replace=False)]
            ds.n_batches = {k: len(ds.batch_idxs[k]) for k in _SPLITS}
            ratios.append(ds.n_batches['train'])
        self.batch_ratios = np.array(ratios) / np.sum(ratios)
        self.n_tot_batches = {k: int(np.sum([ds.n_batches[k] for ds in self.datasets])) for k in _SPLITS}
        self._order_rng = np.random.RandomState(rng_seed)       # same stream on every rank
        self._streams = {}                                      # dtype -> running epoch
        self._lock = threading.Lock()
        self._copy_stream = None
        self._div255 = None
        self._group_size = 1                                    # trials per next_batch call (Multi: sessions per batch)
        for k in _SPLITS:
            self._streams[k] = None

    def __len__(self):
        return self.n_datasets

    def __str__(self):
        s = 'Generator contains %i prefetched session(s):\n' % self.n_datasets
        for ds in self.datasets:
            s += '%s\n    signals: %s\n' % (ds.sess_str, ds.signals)
        return s

    # -- epoch order ---------------------------------------------------------------------------
    def _plan(self, dtype):
        """(session, trial) pairs of one pass, drawn the way the reference consumes its iterators:
        session ~ batch_ratios, rejected when exhausted; trials of a session in random order."""
        rng = self._order_rng
        left = [list(rng.permutation(ds.batch_idxs[dtype])) for ds in self.datasets]
        todo = sum(len(v) for v in left)
        plan = []
        while todo:
            s = int(rng.choice(self.n_datasets, p=self.batch_ratios))
            if not left[s]:
                continue                      # every session has train trials, so every ratio is > 0
            plan.append([(s, int(left[s].pop(0)))])
            todo -= 1
        return plan

    # -- worker --------------------------------------------------------------------------------
    def _frame_range(self, T):
        if not self.shard_frames:
            return 0, T
        from .. import parallel
        return parallel.shard_range(T)

    def _load_host(self, s, idx):
        """Host arrays of one trial: ({signal: array}, (first frame, trial length))."""
        ds = self.datasets[s]
        T = ds.source.trial_length(idx)
        lo, hi = self._frame_range(T)
        out = OrderedDict()
        for sig in ds.signals:
            a = ds.source.load(sig, idx, lo, hi)
            if sig in _INT_SIGNALS:
                a = np.asarray(a, dtype=np.int64)
            elif sig == 'images' and a.dtype == np.uint8:
                pass                                             # scaled after the copy
            elif sig == 'images' and a.dtype != np.float32:
                a = np.asarray(a, dtype=np.float32) / 255        # data_generator.py:258-263
            else:
                a = np.asarray(a, dtype=np.float32)
            tf = self.transforms.get(sig)
            if tf is not None:
                if a.dtype == np.uint8:
                    a = a.astype(np.float32) / 255
                a = np.asarray(tf(a))
            out[sig] = a
        return out, (lo, T)

    def _to_device(self, host, slot):
        """Pinned staging + asynchronous copies on the side stream; returns device tensors and the event."""
        import torch
        dev = torch.device(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        if slot.event is not None:
            slot.event.synchronize()                             # the slot's previous copies have drained
        out = OrderedDict()
        with torch.cuda.stream(self._copy_stream):
            for sig, a in host.items():
                a = np.ascontiguousarray(a)
                pin = slot.buffer(sig, a.nbytes)
                view = pin.numpy().view(a.dtype).reshape(a.shape)
                np.copyto(view, a)
                t = torch.from_numpy(view).to(dev, non_blocking=True)
                if sig == 'images' and t.dtype == torch.uint8 and not self.raw_uint8:
                    if self._div255 is None:
                        self._div255 = torch.full((), 255.0, dtype=torch.float32, device=dev)
                    t = torch.div(t.to(torch.float32), self._div255)     # IEEE division, as numpy's / 255
                out[sig] = t.unsqueeze(0)
            slot.event = torch.cuda.Event()
            slot.event.record(self._copy_stream)
        return out, slot.event

    def _finish_host(self, host):
        import torch
        out = OrderedDict()
        for sig, a in host.items():
            if sig == 'images' and a.dtype == np.uint8 and not self.raw_uint8:
                a = a.astype(np.float32) / 255
            out[sig] = a[None] if self.as_numpy else torch.from_numpy(np.ascontiguousarray(a)).unsqueeze(0)
        return out

    def _worker(self, plan, q, stop):
        try:
            on_gpu = (not self.as_numpy) and str(self.device).startswith('cuda')
            slots = [_PinnedSlot() for _ in range((self.depth + 2) * self._group_size)] if on_gpu else None
            n_slot = 0
            for group in plan:                                   # a group = the trials served by one call
                if stop.is_set():
                    return
                samples, sessions, events = [], [], []
                for s, idx in group:
                    host, shard = self._load_host(s, idx)
                    if on_gpu:
                        sample, event = self._to_device(host, slots[n_slot % len(slots)])
                        n_slot += 1
                        events.append(event)
                    else:
                        sample = self._finish_host(host)
                    sample['batch_idx'] = idx if self.as_numpy else _as_index(idx)
                    if self.shard_frames:
                        sample['shard'] = shard
                    samples.append(sample)
                    sessions.append(s)
                while not stop.is_set():
                    try:
                        q.put((samples, sessions, events), timeout=0.1)
                        break
                    except queue.Full:
                        continue
            q.put(None)
        except BaseException as e:                               # surfaces in next_batch
            q.put(e)

    # -- reference protocol --------------------------------------------------------------------
    def reset_iterators(self, dtype):
        """Start a fresh pass over ``dtype`` ('train' | 'val' | 'test' | 'all')."""
        for k in (_SPLITS if dtype == 'all' else (dtype,)):
            self._start(k, self._plan(k))

    def _start(self, k, plan):
        """(Re)start the worker that serves ``plan`` under stream key ``k``."""
        self._stop(k)
        q = queue.Queue(maxsize=self.depth)
        stop = threading.Event()
        th = threading.Thread(target=self._worker, args=(plan, q, stop), daemon=True)
        self._streams[k] = (q, stop, th)
        th.start()

    def _stop(self, k):
        st = self._streams.get(k)
        if st is not None:
            q, stop, th = st
            stop.set()
            while th.is_alive():
                try:
                    q.get_nowait()
                except queue.Empty:
                    pass
                th.join(timeout=0.05)
            self._streams[k] = None

    def next_batch(self, dtype):
        """(sample, dataset): the next trial of ``dtype``; a new pass starts when the previous one is spent
        or none has been started (the reference builds its iterators in the constructor)."""
        samples, sessions = self._next_group(dtype)
        return samples[0], sessions[0]

    def _next_group(self, key):
        """The next planned group of ``key``: ([samples], [sessions]); copies are ordered before the caller's
        stream.  Raises StopIteration when the pass is spent."""
        if self._streams.get(key) is None:
            self.reset_iterators(key)
        q = self._streams[key][0]
        item = q.get()
        if item is None:
            self._streams[key][2].join()
            self._streams[key] = None
            raise StopIteration('no %s trials left: call reset_iterators' % key)
        if isinstance(item, BaseException):
            self._streams[key] = None
            raise item
        samples, sessions, events = item
        if events:
            import torch
            cur = torch.cuda.current_stream()
            for event in events:
                cur.wait_event(event)
            for sample in samples:
                for v in sample.values():
                    if torch.is_tensor(v) and v.is_cuda:
                        v.record_stream(cur)
        return samples, sessions

    def close(self):
        for k in list(self._streams):
            self._stop(k)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PrefetchSessionsGeneratorMulti(PrefetchSessionsGenerator):
    """Drop-in for the reference's ``ConcatSessionsGeneratorMulti`` (data_generator.py:636-800): a training call
    serves one trial from each of ``n_sessions_per_batch`` DIFFERENT sessions (the MSPS-VAE's triplet loss needs
    them), as ``([samples], [sessions])``; validation / test calls serve single trials."""

    def __init__(self, sources, n_sessions_per_batch=2, **kwargs):
        if n_sessions_per_batch > 4:
            raise NotImplementedError          # the triplet loss covers 2-4 sessions
        if kwargs.get('as_numpy', False):
            raise NotImplementedError
        super().__init__(sources, **kwargs)
        self.n_sessions_per_batch = n_sessions_per_batch
        self._group_size = n_sessions_per_batch
        self._single_train = self.n_tot_batches['train']
        self.n_tot_batches['train'] = int(self.n_tot_batches['train'] / n_sessions_per_batch)

    def __str__(self):
        return 'Multi' + super().__str__()

    def _plan_groups(self):
        """Groups drawn like the reference: per slot a session ~ the renormalised ratios of the sessions not yet
        in the group; a session found exhausted is dropped for this group; the pass ends when fewer sessions
        than open slots remain (trials already drawn for the unfinished group are lost, as in the reference)."""
        rng = self._order_rng
        left = [list(rng.permutation(ds.batch_idxs['train'])) for ds in self.datasets]
        plan = []
        while True:
            ratios = np.copy(self.batch_ratios)
            group = []
            for slot in range(self.n_sessions_per_batch):
                while True:
                    if np.sum(ratios > 0) < self.n_sessions_per_batch - slot:
                        return plan
                    s = int(rng.choice(self.n_datasets, p=ratios))
                    ratios[s] = 0
                    if np.sum(ratios) > 0:
                        ratios = ratios / np.sum(ratios)
                    if left[s]:
                        group.append((s, int(left[s].pop(0))))
                        break
            plan.append(group)

    def reset_iterators(self, dtype):
        super().reset_iterators(dtype)                       # single-trial passes (val / test / train singles)
        if dtype in ('train', 'all'):
            self._multi_spent = False
            self._start('train#multi', self._plan_groups())

    def next_batch(self, dtype, return_multiple=True):
        if dtype == 'train' and return_multiple:
            if getattr(self, '_multi_spent', False):
                return None, None                            # until reset_iterators, like the reference
            if self._streams.get('train#multi') is None:
                self._start('train#multi', self._plan_groups())
            try:
                return self._next_group('train#multi')
            except StopIteration:
                self._multi_spent = True
                return None, None                            # what the reference returns when sessions run out
        return super().next_batch(dtype)


def _as_index(idx):
    import torch
    return torch.tensor([idx])


# the reference's name, for ``from behavenet.data.data_generator import ConcatSessionsGenerator`` swaps
ConcatSessionsGenerator = PrefetchSessionsGenerator
ConcatSessionsGeneratorMulti = PrefetchSessionsGeneratorMulti
