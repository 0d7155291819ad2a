"""The shipped copy of the reference (oracle/_ref, made by oracle/make_ref.py) is byte-for-byte what was copied:
every file still has the sha256 recorded in its manifest, and it matches /root/reference where that exists."""
import filecmp
import os

import pytest

from conftest import ROOT


def test_ref_copy_is_unmodified():
    from oracle import make_ref
    if not os.path.isdir(os.path.join(make_ref.DST, "seistorch")):
        pytest.skip("oracle/_ref not built (run python -m oracle.make_ref where /root/reference exists)")
    assert make_ref.verify()
    if os.path.isdir(os.path.join(make_ref.SRC, "seistorch")):
        cmp = filecmp.dircmp(os.path.join(make_ref.SRC, "seistorch"), os.path.join(make_ref.DST, "seistorch"),
                             ignore=["__pycache__"])
        assert not cmp.diff_files and not cmp.left_only and not cmp.right_only


def test_nothing_in_the_product_imports_the_oracle():
    import re
    pkg = os.path.join(ROOT, "seistorch_b200")
    for root, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(root, f)
