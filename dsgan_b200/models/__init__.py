"""Model registry with the reference's contract (DSGAN/models/__init__.py:4-37): `--model X` resolves to the class
`XModel` (case-insensitive, underscores dropped) in `models/X_model.py`, which must subclass BaseModel."""
import importlib

from .base_model import BaseModel


def find_model_using_name(model_name):
    modellib = importlib.import_module("%s.%s_model" % (__name__, model_name))
    target = model_name.replace("_", "") + "model"
    for name, cls in vars(modellib).items():
        if name.lower() == target.lower() and isinstance(cls, type) and issubclass(cls, BaseModel):
            return cls
    raise NotImplementedError("In %s_model.py, there should be a subclass of BaseModel with class name that matches "
                              "%s in lowercase." % (model_name, target))


def get_option_setter(model_name):
    return find_model_using_name(model_name).modify_commandline_options


def create_model(opt):
    instance = find_model_using_name(opt.model)()
    instance.initialize(opt)
    print("model [%s] was created" % instance.name())
    return instance
