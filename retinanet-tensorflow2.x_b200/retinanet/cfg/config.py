"""JSON config -> attribute dict (reference: retinanet/cfg/config.py:8-21, which wraps easydict.EasyDict).

Both attribute access (`params.inference.mode`, model/builder.py:181) and item access
(`inference_params['max_detections']`, onnx_utils.py:41) are used by the reference, so both work here.  The
reference's JSON files load unmodified.
"""
import json


class AttrDict(dict):
    """Minimal EasyDict: nested dicts (also inside lists) become AttrDicts; attributes and items are the same."""

    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = dict(d or {}, **kwargs)
        for k, v in d.items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __delattr__(self, k):
        try:
            del self[k]
        except KeyError:
            raise AttributeError(k)

    def update(self, *a, **kw):
        for k, v in dict(*a, **kw).items():
            self[k] = v


class Config:
    def __init__(self, path):
        self.path = path
        self._load()

    def _load(self):
        with open(self.path, 'r') as fp:
            self._params = AttrDict(json.load(fp))

    @property
    def params(self):
        return self._params
