"""Device time of the linear autoencoder's training call (AE.loss, model_type='linear') on the C2 frame shape:
256 frames of 128 x 128, 12 latents.  Algorithmic traffic: the frames are streamed three times (encode, decode +
loss, encoder weight gradient) = 3 x 16.8 MB; printed next to the measured copy bandwidth of MEASURED_PEAKS.json."""
import copy, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import cae_oracle as co
from behavenet_b200.models import AE

peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
for n, L in ((256, 12), (2048, 12), (2048, 40)):
    hp = co.make_linear_hparams(1, 128, 128, L)
    model = AE(copy.deepcopy(hp)); model.load_state_dict(co.init_linear_state_dict(hp, seed=0)); model.cuda()
    x = torch.rand(n, 1, 128, 128, device='cuda')
    data = {'images': x[None]}
    for _ in range(5):
        model.loss(data)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        model.loss(data)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gb = 3 * x.numel() * 4 / 1e9
    print('linear AE loss fwd+bwd: n=%d L=%d  %.3f ms per call  %.0f frames/s  %.0f GB/s algorithmic (%.2f of %.0f GB/s)' % (
        n, L, ms, n / ms * 1e3, gb / ms * 1e3, gb / ms * 1e3 / peak, peak))
