"""Encode-only path: latents for the ARHMM (reference ``behavenet/fitting/eval.py:6-118`` export_latents).

``encode_trials`` is the B200 form of the reference loop ``model.encoding(data['images'][0])`` per
trial (eval.py:74-98): the encoder is frame-independent, so whole groups of trials go through ONE
launch sequence, uint8 video crosses PCIe as bytes and is scaled by 1/255 inside the first layer's
patch loader (data_generator.py:258-263 does it on the host), and the latents stay in HBM so that ``HMM.stage_device`` can feed the E-step without a
host round trip (config C5).  ``export_latents`` keeps the reference's call signature and pickle
format for callers that want files.
"""

import numpy as np


def _latents_of(model, frames, labels_2d=None, dataset=None):
    """Latents of a batch of frames the way the reference's export loop forms them (eval.py:50-98): the one-hot
    label images join the frames of a conditional encoder (51-55, 70-71, 87-88), PS-VAE / MSPS-VAE concatenate
    their subspaces (75-76; vaes.py:1255-1266), and the latents of an AE with matrix subspace projection are
    exported in the [labels | rest] basis ``model.U`` (79-81, 95-97)."""
    import torch
    mc = model.hparams.get('model_class')
    if mc == 'cond-ae' and model.hparams.get('conditional_encoder', False):
        if labels_2d is None:
            raise KeyError("a conditional encoder needs data['labels_sc'] (one-hot label images) next to the frames")
        frames = torch.cat((frames.float() if frames.dtype != torch.float32 else frames,
                            labels_2d.to(frames.device, torch.float32)), dim=1)
    out = model.encoding(frames.contiguous(), dataset=dataset)
    if mc == 'ps-vae':
        return torch.cat([out[0], out[1]], dim=1)
    if mc == 'msps-vae':
        return torch.cat([out[0], out[1], out[2]], dim=1)
    if mc == 'cond-ae-msp':
        return model.U(out[0])
    return out[0]


def _scale_u8(t):
    """float32(v) / 255 with an IEEE division (a python-scalar divisor may become a reciprocal multiply)."""
    import torch
    return torch.div(t.to(torch.float32), torch.full((), 255.0, dtype=torch.float32, device=t.device))


def _u8_loader_ok(model):
    """The first layer's uint8 loader (bn_cae_encode_u8) covers <= 4 channels, kernel 5, stride 2."""
    hp = model.hparams
    if hp.get('model_type', 'conv') != 'conv' or hp.get('ae_padding_type', 'same') != 'same':
        return False
    return (hp['ae_input_dim'][0] <= 4 and hp['ae_encoding_kernel_size'][0] == 5 and
            hp['ae_encoding_stride_size'][0] == 2 and hp['ae_encoding_n_channels'][0] % 32 == 0)


def encode_trials(model, trials, frames_per_launch=4096, device=None):
    """Latent means of a list of trials.

    trials: list of (T_i, C, H, W) arrays / tensors, uint8 (0..255) or float32 (0..1), host or device.
    Returns (latents, lengths): a (sum T_i, n_latents) CUDA float32 tensor and the list of T_i.

    Trials are grouped into launches of up to ``frames_per_launch`` frames.  Host trials are copied on a side
    stream one group AHEAD of the encoder (from pinned memory the copy of group k + 1 runs under the kernels of
    group k; pageable memory still works, without the overlap).
    """
    import torch
    if device is None:
        device = next(model.parameters()).device
    device = torch.device(device)
    lengths = [int(t.shape[0]) for t in trials]
    total = int(sum(lengths))
    L = int(model.hparams['n_ae_latents'])
    lat = torch.empty(total, L, dtype=torch.float32, device=device)
    groups, group, gsize = [], [], 0
    for t in trials:
        if gsize and gsize + t.shape[0] > frames_per_launch:
            groups.append(group)
            group, gsize = [], 0
        group.append(t)
        gsize += int(t.shape[0])
    if group:
        groups.append(group)
    on_gpu = device.type == 'cuda'
    compute = torch.cuda.current_stream(device) if on_gpu else None
    side = _side_stream(device) if on_gpu else None

    def stage(group):
        """Enqueue the host -> device copies of one group on the side stream; returns (frames, event)."""
        if not on_gpu:
            return _assemble(model, group, device), None
        side.wait_stream(compute)            # (the group's device-resident trials were produced on `compute`)
        with torch.cuda.stream(side):
            x = _assemble(model, group, device)
            ev = torch.cuda.Event()
            ev.record(side)
        return x, ev

    was_training = model.training
    model.eval()
    with torch.no_grad():
        o = 0
        staged = stage(groups[0]) if groups else None
        for k in range(len(groups)):
            x, ev = staged
            staged = stage(groups[k + 1]) if k + 1 < len(groups) else None
            if ev is not None:
                compute.wait_event(ev)
                x.record_stream(compute)
            lat[o:o + x.shape[0]] = _latents_of(model, x)
            o += x.shape[0]
    model.train(was_training)
    return lat, lengths


_SIDE_STREAMS = {}


def _side_stream(device):
    """ONE copy stream per device for the life of the process: the caching allocator keeps a pool per stream, so a
    fresh stream per call would re-allocate every staging buffer (measured: C5 end to end 563 k -> 376 k frames/s)."""
    import torch
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device)
    return _SIDE_STREAMS[key]


def _assemble(model, group, device):
    """One launch's frames on the device: uint8 stays uint8 when the first layer can read bytes."""
    import torch
    parts = []
    for t in group:
        t = t if torch.is_tensor(t) else torch.from_numpy(np.ascontiguousarray(t))
        t = t.to(device, non_blocking=True)
        parts.append(t if t.dtype == torch.uint8 else t.float())
    if len({q.dtype for q in parts}) > 1:
        parts = [_scale_u8(q) if q.dtype == torch.uint8 else q for q in parts]
    x = parts[0] if len(parts) == 1 else torch.cat(parts, 0)
    if x.dtype == torch.uint8 and not _u8_loader_ok(model):
        x = _scale_u8(x)
    return x


def export_latents(data_generator, model, filename=None):
    """Reference-compatible export (eval.py:6-118): one pickle per dataset with
    ``{'latents': [per-trial arrays, gap trials empty], 'trials': dataset.batch_idxs}``.  The multi-session
    PS-VAE rebuilds a single-session generator over every trial first (eval.py:33-35 -> vaes.py:1202-1273)."""
    import os
    import pickle
    import torch
    if model.hparams.get('model_class') == 'msps-vae' and getattr(data_generator, 'n_sessions_per_batch', 1) > 1:
        return model.export_latents(data_generator, filename=filename)
    model.eval()
    cond_enc = model.hparams.get('model_class') == 'cond-ae' and model.hparams.get('conditional_encoder', False)
    latents = [[np.array([]) for _ in range(ds.n_trials)] for ds in data_generator.datasets]
    with torch.no_grad():
        for dtype in ['train', 'val', 'test']:
            data_generator.reset_iterators(dtype)
            for _ in range(data_generator.n_tot_batches[dtype]):
                data, sess = data_generator.next_batch(dtype)
                y = data['images'][0]
                idx = data['batch_idx'].item() if hasattr(data['batch_idx'], 'item') else int(data['batch_idx'])
                # (the reference walks > 200-frame trials in chunks to fit its GPU; one launch sequence here)
                latents[sess][idx] = _latents_of(model, y, data['labels_sc'][0] if cond_enc else None,
                                                 dataset=sess).cpu().numpy()
    filenames = []
    for sess, ds in enumerate(data_generator.datasets):
        if filename is None:
            sess_id = '%s_%s_%s_%s_latents.pkl' % (ds.lab, ds.expt, ds.animal, ds.session)
            fname = os.path.join(model.hparams['expt_dir'], 'version_%i' % model.version, sess_id)
        else:
            fname = filename
        print('saving latents %i of %i:\n%s' % (sess + 1, data_generator.n_datasets, fname))
        with open(fname, 'wb') as f:
            pickle.dump({'latents': latents[sess], 'trials': ds.batch_idxs}, f)
        filenames.append(fname)
    return filenames


def export_states(hparams, data_generator, model, filename=None):
    """Most likely ARHMM state sequence of every trial, one pickle per dataset (reference eval.py:120-181):
    ``{'states': [per-trial int arrays, gap trials empty], 'trials': dataset.batch_idxs}``.

    The reference calls ``hmm.most_likely_states`` once per trial (eval.py:167); here the trials of a split are
    collected first and decoded by ONE Viterbi launch (``HMM.most_likely_states_batch``) when the model offers
    it -- any object with only ``most_likely_states`` (ssm's interface) is served trial by trial."""
    import os
    import pickle
    states = [[np.array([]) for _ in range(ds.n_trials)] for ds in data_generator.datasets]
    key = 'labels' if hparams['model_class'].find('label') > -1 else 'ae_latents'
    for dtype in ['train', 'val', 'test']:
        data_generator.reset_iterators(dtype)
        where, trials = [], []
        for _ in range(data_generator.n_tot_batches[dtype]):
            data, sess = data_generator.next_batch(dtype)
            y = data[key][0]
            y = y[0] if isinstance(y, (list, tuple)) or (hasattr(y, 'ndim') and y.ndim == 3) else y
            y = y.detach().cpu().numpy() if hasattr(y, 'detach') else np.asarray(y)
            idx = data['batch_idx'].item() if hasattr(data['batch_idx'], 'item') else int(data['batch_idx'])
            where.append((sess, idx))
            trials.append(np.asarray(y, dtype=np.float32))
        if not trials:
            continue
        if hasattr(model, 'most_likely_states_batch'):
            decoded = model.most_likely_states_batch(trials)
        else:
            decoded = [model.most_likely_states(y) for y in trials]
        for (sess, idx), z in zip(where, decoded):
            states[sess][idx] = z
    filenames = []
    for sess, ds in enumerate(data_generator.datasets):
        if filename is None:
            sess_id = '%s_%s_%s_%s_states.pkl' % (ds.lab, ds.expt, ds.animal, ds.session)
            fname = os.path.join(hparams['expt_dir'], 'version_%i' % hparams['version'], sess_id)
        else:
            fname = filename
        print('saving states %i of %i:\n%s' % (sess + 1, data_generator.n_datasets, fname))
        with open(fname, 'wb') as f:
            pickle.dump({'states': states[sess], 'trials': ds.batch_idxs}, f)
        filenames.append(fname)
    return filenames


# model_class -> (position of the latent mean in forward()'s tuple, extra forward kwargs it understands)
_FORWARD_LAYOUT = {
    'ae': (1, ()),
    'cond-ae-msp': (1, ()),
    'vae': (1, ('use_mean',)),
    'beta-tcvae': (1, ('use_mean',)),
    'ps-vae': (2, ('use_mean',)),
    'msps-vae': (2, ('use_mean',)),
    'cond-ae': (1, ('labels', 'labels_2d')),
    'cond-vae': (1, ('labels', 'labels_2d')),
}


def get_reconstruction(model, inputs, dataset=None, return_latents=False, labels=None, labels_2d=None,
                       apply_inverse_transform=True, use_mean=True):
    """Reconstructed frames from frames (4-D input: full forward pass) or from latents (2-D input: decoder
    only), as numpy arrays (reference eval.py:284-376).  Latent inputs of the label-structured models are first
    mapped back to the encoder's basis (``get_inverse_transformed_latents``) unless ``apply_inverse_transform``
    is off; conditional models get their labels appended."""
    import torch
    model.eval()
    if not isinstance(inputs, torch.Tensor):
        inputs = torch.Tensor(inputs).to(model.hparams['device'])
    mc = model.hparams['model_class']
    with torch.no_grad():
        if inputs.dim() != 2:
            if mc not in _FORWARD_LAYOUT:
                raise ValueError('Invalid model class %s' % mc)
            pos, extra = _FORWARD_LAYOUT[mc]
            offered = {'use_mean': use_mean, 'labels': labels, 'labels_2d': labels_2d}
            out = model(inputs, dataset=dataset, **{k: offered[k] for k in extra})
            ims_recon, latents = out[0], out[pos]
        else:
            if mc in ('cond-ae', 'cond-vae'):
                inputs = torch.cat((inputs, labels), dim=1)
            elif mc in ('cond-ae-msp', 'ps-vae', 'msps-vae') and apply_inverse_transform:
                inputs = model.get_inverse_transformed_latents(inputs, as_numpy=False)
            ims_recon = model.decoding(inputs.contiguous(), None, None, dataset=None)
            latents = inputs
    ims_recon = ims_recon.cpu().detach().numpy()
    latents = latents.cpu().detach().numpy()
    return (ims_recon, latents) if return_latents else ims_recon
