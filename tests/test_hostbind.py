"""Host-side helper that pins a rank next to its GPU (pldepth_b200/hostbind.py): parsing only, no GPU."""
from pldepth_b200 import hostbind


def test_cpulist_parsing():
    assert hostbind._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostbind._parse_cpulist("") == set()
    assert hostbind._parse_cpulist("5") == {5}
