"""pytest configuration: markers, shared fixtures and paths."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (REPO, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

FIXTURES = os.path.join(REPO, "fixtures")
GOLDEN = os.path.join(HERE, "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def model_files():
    names = [f for f in os.listdir(os.path.join(FIXTURES, "profile_HMMs")) if f.endswith(".hmm")]
    return sorted(names, key=lambda s: int(s.split(".")[0]))


def hmm_path(name: str) -> str:
    return os.path.join(FIXTURES, "profile_HMMs", name)


def fasta_path(name: str) -> str:
    return os.path.join(FIXTURES, "FASTA_files", name)


def load_golden(name: str):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle, build_oracle

    build_oracle()
    return Oracle()


@pytest.fixture(scope="session")
def reflib():
    from oracle_lib import RefLib

    if not RefLib.available():
        pytest.skip("oracle/_ref/libmsv_ref.so not built (needs /root/reference at build time)")
    return RefLib()


@pytest.fixture(scope="session")
def golden_scores():
    return load_golden("msv_scores.json")


@pytest.fixture(scope="session")
def golden_tables():
    return load_golden("model_tables.json")


@pytest.fixture(scope="session")
def golden_readers():
    return load_golden("readers.json")
