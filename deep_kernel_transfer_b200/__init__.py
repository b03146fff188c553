"""dktb200 -- B200-native implementation of Deep Kernel Transfer's per-episode GP inference path
(backbone forward/backward -> Gram -> Cholesky(K + sigma^2 I) -> marginal likelihood / predictive mean)
behind the reference's ``methods/`` + ``backbone.py`` Python API.  See DESIGN.md."""
__version__ = "0.1.0"
