"""The documents cite files, tests and symbols of this repository: keep the citations alive."""
import os
import re

import pytest

from conftest import ROOT

DOCS = ["DESIGN.md", "INTEGRATION.md", "README.md"]
PREFIXES = ("profiles/", "tests/", "tools/", "fitnoise_b200/", "include/", "oracle/", "csrc/")


def _cited_paths(text):
    for token in re.findall(r"`([^`\n]+)`", text):
        token = token.strip()
        if not token.startswith(PREFIXES) or any(ch in token for ch in "*…<>{} "):
            continue
        yield token


@pytest.mark.parametrize("doc", DOCS)
def test_cited_files_exist(doc):
    text = open(os.path.join(ROOT, doc)).read()
    missing = []
    for token in _cited_paths(text):
        path, _, test_name = token.partition("::")
        if path.startswith("csrc/"):
            path = "fitnoise_b200/" + path
        full = os.path.join(ROOT, path)
        if path.startswith("fitnoise_b200/_lib/") or path.rstrip("/").endswith("_ref"):
            continue                                  # build products / git-ignored locations
        if not os.path.exists(full):
            # profiles are sometimes cited by a prefix ("profiles/r01_bench_n" + "*"): accept a prefix match
            folder, stem = os.path.split(full)
            if os.path.isdir(folder) and any(f.startswith(stem) for f in os.listdir(folder)):
                continue
            missing.append(token)
            continue
        if test_name and os.path.isfile(full):
            if not re.search(r"def %s\b" % re.escape(test_name.split("[")[0]), open(full).read()):
                missing.append(token)
    assert not missing, "%s cites things that do not exist: %s" % (doc, missing)


def test_header_symbols_cited_in_integration_exist():
    header = open(os.path.join(ROOT, "include", "fitnoise_b200.h")).read()
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    cited = set(re.findall(r"\b(fn_[a-z_]+)\(", text))
    declared = set(re.findall(r"\b(fn_[a-z_]+)\(", header))
    assert cited and cited <= declared, sorted(cited - declared)
