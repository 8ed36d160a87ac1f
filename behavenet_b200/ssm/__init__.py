from behavenet_b200.ssm.hmm import HMM

__all__ = ['HMM']
