"""Loss terms with the reference's names and signatures (reference ``behavenet/fitting/losses.py``).

The fused training paths (``AE.loss``, ``VAE.loss``, ``BetaTCVAE.loss``, ``PSVAE.loss``) evaluate these terms and
their gradients inside the CUDA kernels (decoder epilogue, ``bn_psvae_latent``); the functions here are the same
quantities as plain torch expressions on whatever tensors they are given, for the model classes whose loss has
terms outside that fused pass (conditional models, AEMSP, MSPS-VAE) and for analysis code written against
``behavenet.fitting.losses``.
"""

import numpy as np
import torch

LN2PI = float(np.log(2 * np.pi))


def mse(y_pred, y_true, masks=None):
    """Mean over ALL elements of (y_pred - y_true)^2 [* masks] (losses.py:36-59)."""
    d = (y_pred - y_true) ** 2
    return torch.mean(d * masks) if masks is not None else torch.mean(d)


def gaussian_ll(y_pred, y_mean, masks=None, std=1):
    """Log-likelihood under N(y_mean, std^2 I): summed over dims, averaged over the batch (losses.py:62-96)."""
    n_dims = int(np.prod(y_pred.shape[1:]))
    d = (y_pred - y_mean) ** 2
    if masks is not None:
        d = d * masks
    per_frame = d.reshape(d.shape[0], -1).sum(1)
    const = float((0.5 * LN2PI + 0.5 * np.log(std ** 2)) * n_dims)          # python floats: no numpy-scalar (op) tensor
    return torch.mean(-const - float(0.5 / (std ** 2)) * per_frame)


def gaussian_ll_to_mse(ll, n_dims, gaussian_std=1, mse_std=1):
    """Strip the Gaussian constants from a log-likelihood value and rescale it to an MSE (losses.py:99-127; like the
    reference, not meaningful when the likelihood was masked)."""
    v = np.copy(ll) + (0.5 * LN2PI + 0.5 * np.log(gaussian_std ** 2)) * n_dims
    return v * (-(gaussian_std ** 2) / 0.5) / n_dims / (mse_std ** 2)


def kl_div_to_std_normal(mu, logvar):
    """KL(N(mu, diag exp(logvar)) || N(0, I)), averaged over the batch (losses.py:130-147)."""
    return torch.mean(0.5 * torch.sum(logvar.exp() - logvar + mu ** 2 - 1, dim=1))


def _gaussian_log_density_unsummed(z, mu, logvar):
    """log N(z; mu, exp(logvar)) per element (losses.py:354-362)."""
    return -0.5 * ((z - mu) ** 2 * torch.exp(-logvar) + logvar + LN2PI)


def _gaussian_log_density_unsummed_std_normal(z):
    """log N(z; 0, 1) per element (losses.py:365-372)."""
    return -0.5 * (z ** 2 + LN2PI)


def _pairwise(z, mu, logvar):
    """lq[j, i, l] = log q(z_j,l | x_i) and the three reductions every estimator is made of."""
    lq = _gaussian_log_density_unsummed(z[:, None], mu[None, :], logvar[None, :])
    joint = torch.sum(lq, dim=2)
    log_qz = torch.logsumexp(joint, dim=1)                       # log sum_i prod_l q(z_jl | x_i)
    log_qz_product = torch.sum(torch.logsumexp(lq, dim=1), dim=1)     # sum_l log sum_i q(z_jl | x_i)
    return joint, log_qz, log_qz_product


def index_code_mi(z, mu, logvar):
    """Minibatch estimate of the index-code mutual information (losses.py:150-189)."""
    joint, log_qz, _ = _pairwise(z, mu, logvar)
    return torch.mean(torch.diag(joint) - log_qz)


def total_correlation(z, mu, logvar):
    """Minibatch estimate of the total correlation (losses.py:192-241)."""
    _, log_qz, log_qz_product = _pairwise(z, mu, logvar)
    return torch.mean(log_qz - log_qz_product)


def dimension_wise_kl_to_std_normal(z, mu, logvar):
    """Minibatch estimate of the dimension-wise KL to N(0, I) (losses.py:244-281)."""
    _, _, log_qz_product = _pairwise(z, mu, logvar)
    log_pz_product = torch.sum(_gaussian_log_density_unsummed_std_normal(z), dim=1)
    return torch.mean(log_qz_product - log_pz_product)


def decomposed_kl(z, mu, logvar):
    """(index-code MI, total correlation, dimension-wise KL) from one pairwise tensor (losses.py:284-351)."""
    joint, log_qz, log_qz_product = _pairwise(z, mu, logvar)
    log_pz_product = torch.sum(_gaussian_log_density_unsummed_std_normal(z), dim=1)
    return (torch.mean(torch.diag(joint) - log_qz), torch.mean(log_qz - log_qz_product),
            torch.mean(log_qz_product - log_pz_product))


def subspace_overlap(A, B, C=None):
    """mean((U U^T - I)^2) for U = [A; B; (C)]: zero when the stacked rows are orthonormal (losses.py:375-399)."""
    U = torch.cat([A, B] if C is None else [A, B, C], dim=0)
    gram = torch.matmul(U, U.transpose(1, 0))
    return torch.mean((gram - torch.eye(U.shape[0], device=U.device)).pow(2))


def triplet_loss(triplet_loss_obj, z, datasets):
    """Session-separation loss on the background latents (losses.py:402-513) for 2-4 sessions in a batch.

    Every session's frames are shuffled (numpy's global generator, sessions in ascending id order) and dealt
    into 3 * (n - 1) equal chunks: chunks (2k, 2k + 1) are the anchor / positive of the k-th other session,
    whose negative chunk is 2 * (n - 1) + (rank of the anchor session among that session's others).  The sum
    of the triplet terms and of the mean anchor-positive distances is divided by the reference's constant
    (3, 6, 12)."""
    ids = np.unique(datasets)
    n = len(ids)
    if n < 2 or n > 4:
        raise NotImplementedError
    n_chunks = 3 * (n - 1)
    perms = [np.random.permutation(np.where(datasets == i)[0]) for i in ids]
    m = min(len(q) // n_chunks for q in perms)
    chunks = [[q[i::n_chunks][:m] for i in range(n_chunks)] for q in perms]
    loss = 0
    for x in range(n):
        for k, y in enumerate(o for o in range(n) if o != x):
            rank = x if x < y else x - 1
            loss = loss + triplet_loss_obj(z[chunks[x][2 * k]], z[chunks[x][2 * k + 1]],
                                           z[chunks[y][2 * (n - 1) + rank]])
    for x in range(n):
        for k in range(n - 1):
            loss = loss + torch.pairwise_distance(z[chunks[x][2 * k]], z[chunks[x][2 * k + 1]]).mean()
    return loss / {2: 3, 3: 6, 4: 12}[n]


def r2_variance_weighted(y_true, y_pred):
    """``sklearn.metrics.r2_score(y_true, y_pred, multioutput='variance_weighted')`` for 2-D numpy arrays, which is what
    the reference logs as ``label_r2`` after every PS-VAE loss call (vaes.py:709-718).  Same arithmetic and the same
    treatment of constant columns as sklearn's implementation, without its per-call input validation (0.5 ms per
    call on the 512 x 4 labels of config C3 -- 10 % of the whole training step)."""
    import numpy as np
    y_true = np.asarray(y_true)
    y_pred = np.asarray(y_pred)
    if y_true.ndim == 1:
        y_true, y_pred = y_true[:, None], y_pred[:, None]
    if y_true.shape[0] < 2:
        return float('nan')                    # sklearn: undefined (it warns and returns nan)
    num = ((y_true - y_pred) ** 2).sum(axis=0, dtype=np.float64)
    den = ((y_true - np.average(y_true, axis=0)) ** 2).sum(axis=0, dtype=np.float64)
    nz_den, nz_num = den != 0, num != 0
    scores = np.ones(y_true.shape[1])
    valid = nz_den & nz_num
    scores[valid] = 1.0 - num[valid] / den[valid]
    scores[nz_num & ~nz_den] = 0.0
    if not np.any(nz_den):
        return float(np.average(scores))
    return float(np.average(scores, weights=den))
