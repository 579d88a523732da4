"""Loads tests/golden/ (written by oracle/pin_oracle.py from the unmodified reference).

families.json holds, per family of cases, how to derive each case from a base stream
(cut = prefix, flip = one bit flipped, cap = dst capacity sweep, hex = literal streams) and, per
case: class (D defined / A assert-only / U reference-UB), status, bytes written and the FNV-1a
fingerprint of the whole dst capacity (0xA5 pre-fill).  For D and A cases these ARE the
reference's results (pin_oracle.py verified them equal); for U cases they are this
repository's documented choice.
"""
from __future__ import annotations

import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILL = 0xA5


class Golden:
    def __init__(self):
        with open(os.path.join(GOLDEN, "families.json")) as f:
            self.meta = json.load(f)
        self.bases = {}
        for name in self.meta["bases"]:
            with open(os.path.join(GOLDEN, "bases", name + ".deflate"), "rb") as f:
                self.bases[name] = f.read()
        self.families = {f["name"]: f for f in self.meta["families"]}

    def cases(self, fam_name, stride: int = 1):
        """-> list of (index, src bytes, cap)"""
        fam = self.families[fam_name]
        kind = fam["kind"]
        idx = range(0, len(fam["params"]), stride)
        if kind == "hex":
            return [(i, bytes.fromhex(fam["params"][i]), fam["caps"][i]) for i in idx]
        comp = self.bases[fam["base"]]
        if kind == "cut":
            return [(i, comp[:fam["params"][i]], fam["caps"]) for i in idx]
        if kind == "flip":
            out = []
            for i in idx:
                bit = fam["params"][i]
                b = bytearray(comp)
                b[bit >> 3] ^= 1 << (bit & 7)
                out.append((i, bytes(b), fam["caps"]))
            return out
        if kind == "cap":
            return [(i, comp, fam["params"][i]) for i in idx]
        raise ValueError(kind)

    def expected(self, fam_name, i):
        fam = self.families[fam_name]
        return fam["status"][i], fam["written"][i], fam["hash"][i], fam["class"][i]


_g = None


def load() -> Golden:
    global _g
    if _g is None:
        _g = Golden()
    return _g
