"""``get_trainer`` with the reference's lookup rule (ctu/trainers/__init__.py:5-20)."""
import importlib


def get_trainer(opt):
    name = opt.model
    module = importlib.import_module(__name__ + '.' + name + '_trainer')
    target = name.replace('_', '') + 'trainer'
    for cls_name, cls in module.__dict__.items():
        if cls_name.lower() == target.lower() and isinstance(cls, type):
            return cls
    raise ValueError('In {}_trainer.py, there should be a trainer class named {} (case-insensitive).'.format(
        name, target))
