"""Importable alias of the package directory `motion-planning-for-autonomous-driving-with-mpc_b200/`
(a hyphenated directory name cannot be imported directly).  All code lives there."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "motion-planning-for-autonomous-driving-with-mpc_b200")
__path__.append(_PKG_DIR)
PACKAGE_DIR = _PKG_DIR

from .scenarios import load_scenario, scenario_names, perturbed_initial_states, reference_window, make_batch  # noqa: E402,F401
