# ``data.datamgr`` is served from here (GPU episode feeder); the reference's other data modules (feature_loader,
# qmul_loader, dataset ...) stay importable from a checkout further down sys.path.
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
