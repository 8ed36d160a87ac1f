from behavenet_b200.models.aes import AE, ConvAEDecoder, ConvAEEncoder, load_pretrained_ae
from behavenet_b200.models.vaes import PSVAE, ConvAEPSEncoder, reparameterize

__all__ = ['AE', 'ConvAEDecoder', 'ConvAEEncoder', 'load_pretrained_ae', 'PSVAE',
           'ConvAEPSEncoder', 'reparameterize']
