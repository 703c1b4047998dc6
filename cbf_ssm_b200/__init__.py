"""cbf_ssm_b200 -- B200-native sampled-ELBO hot path of CBF-SSM behind the reference's
model / training / dataset interfaces.  The arithmetic lives in ``libcbfssm_b200.so``
(hand-written sm_100a CUDA, C ABI in ``include/cbfssm_b200.h``); there is no CPU path."""
__version__ = "0.1.0"
