"""Model plugin lookup with the reference's rules (ctu/models/__init__.py:10-43): ``--model NAME`` imports
``<this package>.NAME_model`` and picks the ``nn.Module`` subclass called ``NAMEmodel`` (case-insensitive, underscores
dropped); ``get_option_setter`` hands its static ``modify_commandline_options`` to the parser (base_parser.py:142-144)."""
import importlib

import torch


def find_model_using_name(model_name):
    module = importlib.import_module(__name__ + '.' + model_name + '_model')
    target = model_name.replace('_', '') + 'model'
    for name, cls in module.__dict__.items():
        if name.lower() == target.lower() and isinstance(cls, type) and issubclass(cls, torch.nn.Module):
            return cls
    # the reference prints this and calls exit(0); a library should not end the process, so raise instead
    raise ValueError('In %s_model.py, there should be a subclass of torch.nn.Module with class name that matches %s in '
                     'lowercase.' % (model_name, target))


def get_option_setter(model_name):
    return find_model_using_name(model_name).modify_commandline_options


def create_model(opt):
    instance = find_model_using_name(opt.model)(opt)
    print('model [%s] was created' % type(instance).__name__)
    return instance
