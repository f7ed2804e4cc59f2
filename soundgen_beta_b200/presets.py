"""The bundled soundgen() presets (argument values of R/presets.R:156-410, extracted by
scripts/extract_presets.py into data/presets.json) as keyword dicts for soundgen()."""
from __future__ import annotations

import json
import os

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'presets.json')
_cache = None


def _conv(v):
    if isinstance(v, dict):
        if '__na__' in v:
            return None
        return {k: _conv(x) for k, x in v['__list__']}
    if isinstance(v, list):
        return [_conv(x) for x in v]
    return v


_ANCHORS = ('pitchAnchors', 'pitchAnchorsGlobal', 'noiseAnchors', 'mouthAnchors', 'amplAnchors', 'amplAnchorsGlobal')


def _neutral(kw):
    """anchors -> (time[], value[]); formants -> list of (k, 4) arrays (time, freq, amp, width), the forms
    both the CUDA mirror and the oracle take."""
    import numpy as np
    out = {}
    for k, v in kw.items():
        if k in _ANCHORS and isinstance(v, dict):
            v = (np.asarray(v['time'], dtype=np.float64), np.asarray(v['value'], dtype=np.float64))
        elif k in ('formants', 'formantsNoise') and isinstance(v, dict):
            rows = []
            for f in v.values():
                cols = [np.atleast_1d(np.asarray(f[c], dtype=np.float64)) for c in ('time', 'freq', 'amp', 'width')]
                n = max(c.size for c in cols)
                rows.append(np.stack([np.resize(c, n) for c in cols], axis=1))
            v = rows
        out[k] = v
    return out


def load():
    """[(speaker, name, kwargs)] in the order of R/presets.R."""
    global _cache
    if _cache is None:
        raw = json.load(open(_PATH))['presets']
        _cache = [(p['speaker'], p['name'], _neutral({k: _conv(v) for k, v in p['args']})) for p in raw]
    return _cache


def preset(speaker, name):
    for s, n, kw in load():
        if s == speaker and n == name:
            return dict(kw)
    raise KeyError('%s/%s' % (speaker, name))
