"""Per-trial signal transforms served by the data generators (reference ``behavenet/data/transforms.py``).

Host-side numpy, applied by the prefetch worker thread before a trial is staged in pinned memory
(``PrefetchSessionsGenerator(transforms={'labels_sc': MakeOneHot2D(...)})``).  Same class names, constructor
arguments, error behaviour and ``repr`` strings as the reference; ``MakeOneHot2D`` and ``Threshold`` are written
against current numpy (the reference uses the removed ``np.int`` / ``np.float`` aliases, transforms.py:226,355) and
do not modify their input in place.
"""

import numpy as np


class Compose(object):
    """Chain transforms: ``Compose([SelectIdxs(idxs), ZScore()])(signal)`` (transforms.py:10-45)."""

    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, signal):
        for transform in self.transforms:
            signal = transform(signal)
        return signal

    def __repr__(self):
        # the reference's string, backspaces included (transforms.py:40-45)
        return self.__class__.__name__ + '(' + ''.join('{0}, '.format(t) for t in self.transforms) + '\b\b)'


class Transform(object):
    """Base class (transforms.py:48-55)."""

    def __call__(self, *args):
        raise NotImplementedError

    def __repr__(self):
        raise NotImplementedError


class BlockShuffle(Transform):
    """Shuffle the runs of a discrete state sequence, keeping every run intact (transforms.py:58-109); the
    permutation comes from numpy's global generator seeded with ``rng_seed`` on every call."""

    def __init__(self, rng_seed):
        self.rng_seed = rng_seed

    def __call__(self, sample):
        np.random.seed(self.rng_seed)
        n_time = len(sample)
        if np.any(np.isnan(sample)):
            return np.full(n_time, fill_value=np.nan)
        starts = np.flatnonzero(np.diff(sample) != 0) + 1            # first index of every run but the first
        bounds = np.concatenate([[0], starts, [n_time]])
        runs = [np.arange(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1)]
        order = np.random.permutation(len(runs))
        return sample[np.concatenate([runs[i] for i in order])]

    def __repr__(self):
        return str('BlockShuffle(rng_seed=%i)' % self.rng_seed)


class ClipNormalize(Transform):
    """min(signal, clip_val) / clip_val (transforms.py:112-146)."""

    def __init__(self, clip_val):
        if clip_val <= 0:
            raise ValueError('clip value must be positive')
        self.clip_val = clip_val

    def __call__(self, signal):
        return np.minimum(signal, self.clip_val) / self.clip_val

    def __repr__(self):
        return str('ClipNormalize(clip_val=%f)' % self.clip_val)


class MakeOneHot(Transform):
    """(T,) integer states -> (T, max + 1) one-hot rows; 2-D input passes through; NaNs anywhere give an
    all-NaN array (transforms.py:149-183)."""

    def __call__(self, sample):
        if len(sample.shape) == 2:
            return sample
        n_time = len(sample)
        onehot = np.zeros((n_time, int(np.nanmax(sample)) + 1))
        if np.any(np.isnan(sample)):
            onehot[:] = np.nan
        else:
            onehot[np.arange(n_time), sample.astype('int')] = 1
        return onehot

    def __repr__(self):
        return 'MakeOneHot()'


class MakeOneHot2D(Transform):
    """(T, 2 * n) labels [x_0..x_{n-1}, y_0..y_{n-1}] -> (T, n, y_pixels, x_pixels) images with a single one at
    the rounded, clipped (y, x) of each label; NaN coordinates go to 0 (transforms.py:186-248).  These are the
    ``labels_sc`` frames a conditional encoder reads next to the video channels."""

    def __init__(self, y_pixels, x_pixels):
        self.y_pixels = y_pixels
        self.x_pixels = x_pixels

    @staticmethod
    def _pixel(vals, n_pix):
        vals = np.where(np.isnan(vals), -1.0, np.asarray(vals, dtype=np.float64))
        return np.round(np.clip(vals, 0, n_pix - 1)).astype(np.int64)

    def __call__(self, sample):
        time, n2 = sample.shape
        n = n2 // 2
        xs = self._pixel(sample[:, :n], self.x_pixels)
        ys = self._pixel(sample[:, n:], self.y_pixels)
        out = np.zeros((time, n, self.y_pixels, self.x_pixels))
        out[np.arange(time)[:, None], np.arange(n)[None, :], ys, xs] = 1
        return out

    def __repr__(self):
        return str('MakeOneHot2D(y_pixels=%i, x_pixels=%i)' % (self.y_pixels, self.x_pixels))


class MotionEnergy(Transform):
    """|x_t - x_{t-1}| with a zero first row (transforms.py:251-274)."""

    def __call__(self, sample):
        return np.vstack([np.zeros((1, sample.shape[1])), np.abs(np.diff(sample, axis=0))])

    def __repr__(self):
        return 'MotionEnergy()'


class SelectIdxs(Transform):
    """Columns ``idxs`` of a (T, N) signal (transforms.py:277-310)."""

    def __init__(self, idxs, sample_name=''):
        self.sample_name = sample_name
        self.idxs = idxs

    def __call__(self, sample):
        return sample[:, self.idxs]

    def __repr__(self):
        return str('SelectIndxs(idxs=idxs, sample_name=%s)' % self.sample_name)


class Threshold(Transform):
    """Drop neurons whose mean rate (Hz, bins of ``bin_size`` ms) is not above ``threshold`` (transforms.py:313-357)."""

    def __init__(self, threshold, bin_size):
        if bin_size <= 0:
            raise ValueError('bin size must be positive')
        if threshold < 0:
            raise ValueError('threshold must be non-negative')
        self.threshold = threshold
        self.bin_size = bin_size

    def __call__(self, sample):
        rates = np.squeeze(np.mean(sample, axis=0)) / (self.bin_size * 1e-3)
        return sample[:, rates > self.threshold].astype(np.float64)

    def __repr__(self):
        return str('Threshold(threshold=%f, bin_size=%f)' % (self.threshold, self.bin_size))


class ZScore(Transform):
    """Column-wise (x - mean) / std, in place like the reference (transforms.py:360-385)."""

    def __call__(self, sample):
        sample -= np.mean(sample, axis=0)
        sample /= np.std(sample, axis=0)
        return sample

    def __repr__(self):
        return 'ZScore()'
