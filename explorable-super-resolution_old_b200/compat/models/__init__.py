"""`models` package shim.  models.networks comes from this directory; models.SRRaGAN_model, models.base_model and
anything else resolve to the reference's codes/models through the extended package path.  create_model is the
reference package's factory (codes/models/__init__.py:1-10), restated because this __init__ runs in its place."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)


def create_model(opt, *args, **kwargs):
    kind = opt['model']
    if kind != 'srragan':
        raise NotImplementedError('Model [{:s}] not recognized.'.format(kind))
    from .SRRaGAN_model import SRRaGANModel
    model = SRRaGANModel(opt, *args, **kwargs)
    print('Model [{:s}] is created.'.format(model.__class__.__name__))
    return model
