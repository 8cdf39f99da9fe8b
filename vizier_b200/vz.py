"""Resolves the pyvizier data model: the real `vizier.pyvizier` when importable (drop-in inside a
Vizier install), else the bundled stand-in `pyvizier_lite`.  Also exposes the Designer-side
containers (CompletedTrials / ActiveTrials / Prediction)."""
# pylint: disable=wildcard-import,unused-wildcard-import,g-import-not-at-top
try:
  from vizier import pyvizier as _real  # needs compiled protos; absent in offline builds
  from vizier import algorithms as _vza
  from vizier.pyvizier import *  # noqa: F401,F403
  CompletedTrials = _vza.CompletedTrials
  ActiveTrials = _vza.ActiveTrials
  Prediction = _vza.Prediction
  Designer = _vza.Designer
  Predictor = _vza.Predictor
  USING_REAL_VIZIER = True
except Exception:  # pylint: disable=broad-except
  from vizier_b200.pyvizier_lite import *  # noqa: F401,F403
  from vizier_b200.pyvizier_lite import ActiveTrials, CompletedTrials, Designer, Prediction, Predictor  # noqa: F401
  USING_REAL_VIZIER = False
