"""``get_trainer`` with the reference's lookup rule (ctu/trainers/__init__.py:5-20): ``--model NAME`` imports
``<this package>.NAME_trainer`` and returns the class called ``NAMEtrainer`` (case-insensitive, underscores dropped)."""
import importlib


def get_trainer(opt):
    wanted = (opt.model.replace('_', '') + 'trainer').lower()
    namespace = vars(importlib.import_module('%s.%s_trainer' % (__name__, opt.model)))
    hits = [obj for key, obj in namespace.items() if key.lower() == wanted and isinstance(obj, type)]
    if not hits:
        raise ValueError('In {}_trainer.py, there should be a trainer class named {} (case-insensitive).'.format(
            opt.model, wanted))
    return hits[-1]
