from behavenet_b200.models.aes import (AE, AEMSP, ConditionalAE, ConvAEDecoder, ConvAEEncoder, LinearAEDecoder,  # noqa: F401
                                       LinearAEEncoder, load_pretrained_ae)
from behavenet_b200.models.vaes import VAE, BetaTCVAE, ConditionalVAE, MSPSVAE, ConvAEMSPSEncoder, PSVAE, ConvAEPSEncoder, reparameterize

__all__ = ['AE', 'AEMSP', 'ConditionalAE', 'ConvAEDecoder', 'ConvAEEncoder', 'LinearAEDecoder', 'LinearAEEncoder', 'load_pretrained_ae', 'VAE', 'BetaTCVAE', 'ConditionalVAE', 'MSPSVAE', 'ConvAEMSPSEncoder', 'PSVAE',
           'ConvAEPSEncoder', 'reparameterize']
