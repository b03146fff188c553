# ``methods.DKT`` / ``methods.DKT_regression`` / ``methods.meta_template`` are served from here; with a reference checkout
# behind this repo on sys.path its other method modules (baselinetrain, protonet, maml ... -- out of scope) stay importable.
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
