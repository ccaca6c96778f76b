"""TEST INFRASTRUCTURE ONLY -- fixed f'/f'' table standing in for xraydb.

The reference looks up anomalous scattering factors through
xraydb.f1_chantler / f2_chantler (tools/utilities.py:357-358).  xraydb is not
vendored, not version-pinned and not installed here, so the oracle and the
reference shim share this small constant table instead.  Values are
approximate Chantler numbers near 12.7 keV; they are *placeholders*: the hot
path receives f-values as an input array, so parity between the CUDA path and
the oracle does not depend on them ("parity unpinned at the xraydb boundary").
The table ignores `energy` on purpose.
"""

_F1F2 = {
    "H": (0.0, 0.0),
    "C": (0.0049, 0.0023),
    "N": (0.0090, 0.0047),
    "O": (0.0160, 0.0090),
    "F": (0.0240, 0.0140),
    "Si": (0.1100, 0.1000),
    "P": (0.1400, 0.1400),
    "S": (0.1700, 0.2500),
    "Cl": (0.2000, 0.2400),
}


def f1_f2(element, energy=None):
    """(f', f'') for `element`; KeyError for elements not in the table, which
    the reference catches and reports (tools/utilities.py:361-362)."""
    return _F1F2[element]
