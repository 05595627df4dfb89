"""gen.pro / gen.founder (reference geneakit/identify.py:6-53)."""
from . import cgeneakit


def pro(gen):
    return cgeneakit.get_proband_ids(gen)


def founder(gen):
    return cgeneakit.get_founder_ids(gen)
