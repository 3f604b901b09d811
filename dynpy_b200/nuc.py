"""Nuclear constants per isotope: quadrupole moment Q (m^2), spin I, gyromagnetic ratio gamma.

Data of the reference's nuc.py:1-16 (values are part of the drop-in contract: they enter
relaxation() at qrax.py:218-233 and dipolar() at ddrax.py:224-274)."""

nuc = {
    "23Na": {"Z": 11, "Q": 104e-31, "I": 3 / 2, "rot_J": 3.11e27, "gamma": 10.4},
    "7Li": {"Z": 3, "Q": -40e-31, "I": 3 / 2, "rot_J": 9.37e27, "gamma": 0.17},
    "39K": {"Z": 19, "Q": 4.9e-30, "I": 3 / 2, "rot_J": 2.19e27, "gamma": 35.7},
    "85Rb": {"Z": 37, "Q": 2.6e-29, "I": 5 / 2, "rot_J": 1.54e27},
    "133Cs": {"Z": 55, "Q": -3.43e-31, "I": 7 / 2, "rot_J": 0.9e27},
    "127I": {"Z": 53, "Q": -688.22e-31, "I": 5 / 2, "rot_J": None},
    "25Mg": {"Z": 12, "Q": 2.2e-29, "I": 5 / 2, "rot_J": 1.65e27, "gamma": 7.64},
    "43Ca": {"Z": 20, "Q": 2e-29, "I": 7 / 2, "rot_J": 2.0e27},
    "87Sr": {"Z": 38, "Q": 3e-29, "I": 9 / 2, "rot_J": 2.44e27},
    "35Cl": {"Z": 17, "Q": -82e-31, "I": 3 / 2, "rot_J": 7.57e27},
    "17O": {"Z": 8, "Q": -25.6e-31, "I": 5 / 2, "rot_J": None},
    "1H": {"Z": 1, "Q": None, "I": 1 / 2, "rot_J": None, "gamma": 267.5221900e6},
    "2H": {"Z": 1, "Q": 2.8e-31, "I": 1, "rot_J": None},
    "13C": {"Z": 6, "Q": None, "I": 1 / 2, "rot_J": None, "gamma": 67.28e6},
    "14N": {"Z": 7, "Q": 20.4e-31, "I": 1, "rot_J": None},
    "131Xe": {"Z": 54, "Q": -114.6e-31, "I": 3 / 2, "rot_J": None},
    "129Xe": {"Z": 54, "Q": 0, "I": 1 / 2, "rot_J": None, "gamma": -7.441e7},
}

# scipy.constants values the formulas use (CODATA exact SI definitions; mu_0 is the CODATA
# 2018 recommended value scipy >= 1.4 ships).  Kept literal so results do not depend on the
# scipy build of the host.
E_CHARGE = 1.602176634e-19
HBAR = 6.62607015e-34 / (2 * 3.141592653589793)
try:  # follow the installed scipy exactly when it is there (it is what the reference would use)
    from scipy import constants as _c
    E_CHARGE, HBAR, MU_0 = _c.e, _c.hbar, _c.mu_0
except Exception:  # pragma: no cover
    MU_0 = 1.25663706212e-06
