"""test shim: minimal EasyDict (attribute access over nested dicts) for the reference's yaml configs"""


class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if isinstance(value, dict) and not isinstance(value, EasyDict):
            value = EasyDict(value)
        elif isinstance(value, (list, tuple)):
            value = type(value)(EasyDict(x) if isinstance(x, dict) else x for x in value)
        super().__setattr__(name, value)
        super().__setitem__(name, value)

    __setitem__ = __setattr__
