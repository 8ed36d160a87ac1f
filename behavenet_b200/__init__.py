"""behavenet_b200 -- B200-native (sm_100a) hot paths of BehaveNet behind the reference's API.

Hot path 1: convolutional autoencoder / PS-VAE  -> ``behavenet_b200.models`` (AE, PSVAE, ...)
Hot path 2: ARHMM E-step / log-likelihood / Viterbi -> ``behavenet_b200.ssm.HMM``

Everything numerical runs in ``libbehavenet_b200.so`` (hand-written CUDA, C ABI in
``include/behavenet_b200.h``); there is no CPU or eager-PyTorch fallback.
"""

__version__ = '0.1.0'
