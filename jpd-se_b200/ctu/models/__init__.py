"""Model plugin lookup with the reference's rules (ctu/models/__init__.py:10-43): ``--model NAME`` imports
``<this package>.NAME_model`` and picks the ``nn.Module`` subclass called ``NAMEmodel`` (case-insensitive, underscores
dropped); ``get_option_setter`` hands its static ``modify_commandline_options`` to the parser (base_parser.py:142-144)."""
import importlib

from torch import nn


def find_model_using_name(model_name):
    wanted = (model_name.replace('_', '') + 'model').lower()
    namespace = vars(importlib.import_module('%s.%s_model' % (__name__, model_name)))
    hits = [obj for key, obj in namespace.items()
            if key.lower() == wanted and isinstance(obj, type) and issubclass(obj, nn.Module)]
    if not hits:
        # the reference prints this and calls exit(0); a library should not end the process, so raise instead
        raise ValueError('In %s_model.py, there should be a subclass of torch.nn.Module with class name that matches %s '
                         'in lowercase.' % (model_name, wanted))
    return hits[-1]


def get_option_setter(model_name):
    return find_model_using_name(model_name).modify_commandline_options


def create_model(opt):
    model = find_model_using_name(opt.model)(opt)
    print('model [%s] was created' % type(model).__name__)
    return model
