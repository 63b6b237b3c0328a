"""Helpers shared by the golden-vector tests."""
import hashlib

from smallz4_b200 import corpus


def case_input(c):
    return corpus.make(c["kind"], c["size"], c["seed"]).tobytes()


def case_dict(c):
    if not c["dict"]:
        return None
    kind, seed, size = c["dict"]
    return corpus.make(kind, size, seed, offset=1 << 40).tobytes()


def case_id(c):
    d = "" if not c["dict"] else f"-D{c['dict'][2]}"
    return f"{c['kind']}-{c['size']}-L{c['level']}{'-legacy' if c['legacy'] else ''}{d}"


def digest(b):
    return hashlib.sha256(b).hexdigest()
