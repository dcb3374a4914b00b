#!/usr/bin/env python
"""cmp the vendored reference tests with /root/reference/src/python/tests (exit 0 if identical
or if the reference tree is absent, as on the GPU box)."""
import filecmp
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
MINE = os.path.join(HERE, "src", "python", "tests")
REF = "/root/reference/src/python/tests"


def main() -> int:
    if not os.path.isdir(REF):
        print("reference tree not present: nothing to compare")
        return 0
    names = sorted(f for f in os.listdir(REF) if f.endswith(".py"))
    mine = sorted(f for f in os.listdir(MINE) if f.endswith(".py"))
    bad = [n for n in names if n not in mine or not filecmp.cmp(os.path.join(REF, n), os.path.join(MINE, n), shallow=False)]
    bad += [n for n in mine if n not in names]
    print(f"{len(names)} reference test files, {len(bad)} differing: {bad}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
