from behavenet_b200.models.aes import AE, ConvAEDecoder, ConvAEEncoder, load_pretrained_ae
from behavenet_b200.models.vaes import VAE, BetaTCVAE, PSVAE, ConvAEPSEncoder, reparameterize

__all__ = ['AE', 'ConvAEDecoder', 'ConvAEEncoder', 'load_pretrained_ae', 'VAE', 'BetaTCVAE', 'PSVAE',
           'ConvAEPSEncoder', 'reparameterize']
