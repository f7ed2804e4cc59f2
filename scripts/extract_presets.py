"""Extracts the argument values of the bundled soundgen() presets (R/presets.R:156-410 of the reference)
into soundgen_beta_b200/data/presets.json.  Only DATA is taken (the numbers inside the call strings);
run in the build container, where /root/reference is mounted:

    python scripts/extract_presets.py /root/reference/R/presets.R
"""
import json
import os
import re
import sys


class P:
    """Recursive-descent reader of the R literals used in the preset strings:
    numbers, NA, NULL, 'strings', c(...), list(name = value, ...)."""

    def __init__(self, s):
        self.s, self.i = s, 0

    def ws(self):
        while self.i < len(self.s) and self.s[self.i].isspace():
            self.i += 1

    def peek(self):
        self.ws()
        return self.s[self.i] if self.i < len(self.s) else ''

    def expect(self, ch):
        self.ws()
        assert self.s[self.i] == ch, (ch, self.s[self.i:self.i + 30])
        self.i += 1

    def ident(self):
        self.ws()
        m = re.match(r'[A-Za-z_.][A-Za-z0-9_.]*', self.s[self.i:])
        assert m, self.s[self.i:self.i + 30]
        self.i += m.end()
        return m.group(0)

    def value(self):
        c = self.peek()
        if c in '-+0123456789.':
            m = re.match(r'[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?', self.s[self.i:])
            self.i += m.end()
            v = float(m.group(0))
            return int(v) if v == int(v) and '.' not in m.group(0) and 'e' not in m.group(0).lower() else v
        if c in '"\'':
            j = self.s.index(c, self.i + 1)
            v = self.s[self.i + 1:j]
            self.i = j + 1
            return v
        name = self.ident()
        if name == 'NA':
            return {'__na__': True}
        if name == 'NULL':
            return None
        if name in ('TRUE', 'FALSE'):
            return name == 'TRUE'
        assert name in ('c', 'list'), name
        self.expect('(')
        items, names = [], []
        while self.peek() != ')':
            save = self.i
            nm = None
            m = re.match(r'\s*([A-Za-z_.][A-Za-z0-9_.]*)\s*=(?!=)', self.s[self.i:])
            if m:
                nm = m.group(1)
                self.i += m.end()
            else:
                self.i = save
            items.append(self.value())
            names.append(nm)
            if self.peek() == ',':
                self.i += 1
        self.expect(')')
        if name == 'c':
            return items
        if all(n is not None for n in names):
            return {'__list__': [[n, v] for n, v in zip(names, items)]}
        return items

    def call(self):
        assert self.ident() == 'soundgen'
        self.expect('(')
        args = []
        while self.peek() != ')':
            nm = self.ident()
            self.expect('=')
            args.append([nm, self.value()])
            if self.peek() == ',':
                self.i += 1
        return args


def main(path):
    src = open(path).read()
    body = src[src.index('presets = list('):]
    out = []
    speakers = [(m.start(), m.group(1)) for m in re.finditer(r'^  ([A-Za-z0-9_]+) = list\($', body, re.M)]
    for m in re.finditer(r"^    ([A-Za-z0-9_]+) = '(soundgen\(.*?\))'", body, re.M | re.S):   # a call may span lines
        speaker = [nm for pos, nm in speakers if pos < m.start()][-1]
        out.append({'speaker': speaker, 'name': m.group(1), 'args': P(m.group(2)).call()})
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'soundgen_beta_b200', 'data',
                       'presets.json')
    json.dump({'source': 'argument values of the presets in R/presets.R:156-410 (soundgen package data)',
               'presets': out}, open(dst, 'w'), indent=0)
    print(len(out), 'presets ->', dst)
    for p in out:
        print(' ', p['speaker'], p['name'], len(p['args']), 'args')


if __name__ == '__main__':
    main(sys.argv[1])
