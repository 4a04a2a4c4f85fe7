"""Nested-class configuration objects (mirrors reference envs/base/base_config.py:34-56).

A config is a class whose attributes are either plain values or further classes; instantiating
the outermost class turns every nested class into an instance, so `cfg.env.num_envs = 8`
edits that one config object and not the class shared by all users."""
import inspect


class BaseConfig:
    def __init__(self):
        _instantiate_nested(self)


def _instantiate_nested(node):
    for name in dir(node):
        if name == "__class__":
            continue
        member = getattr(node, name)
        if inspect.isclass(member):
            child = member()
            setattr(node, name, child)
            _instantiate_nested(child)


# the reference exposes the walker as a static method too
BaseConfig.init_member_classes = staticmethod(_instantiate_nested)
