"""`DesignerPolicy`: the Pythia policy that wraps a Designer, with the reference's call sequence.

Mirrors vizier/_src/algorithms/policies/designer_policy.py:44-125: every `suggest(request)` builds a
FRESH designer from the study config (`designer_factory(problem)`), fetches COMPLETED and ACTIVE trials
from the policy supporter, calls `designer.update(CompletedTrials, ActiveTrials)` once and then
`designer.suggest(count)`.  Inside a Vizier install the real class is used unchanged (INTEGRATION.md:
`DesignerPolicy(supporter, vizier_b200.VizierGPBandit.from_problem)`, the one-line change at
vizier/_src/service/policy_factory.py:48-53); this stand-in exists so that the drop-in boundary can be
exercised without the service (the reference's pythia package needs generated protos).
"""

from __future__ import annotations

import dataclasses
from typing import Any, Callable, List, Optional, Sequence

from vizier_b200 import vz


@dataclasses.dataclass
class SuggestRequest:
  """pythia.SuggestRequest (vizier/_src/pythia/policy.py): the fields DesignerPolicy reads."""
  study_config: Any          # ProblemStatement
  count: Optional[int] = 1
  study_guid: str = ''
  max_trial_id: int = 0


@dataclasses.dataclass
class SuggestDecision:
  """pythia.SuggestDecision: suggestions (+ metadata delta, unused here)."""
  suggestions: Sequence[Any]
  metadata: Any = None


class InRamPolicySupporter:
  """The subset of `pythia.PolicySupporter` DesignerPolicy uses: `GetTrials(status_matches=...)`
  (vizier/_src/pythia/policy_supporter.py) over an in-memory trial list, plus the convenience the
  reference's local supporter offers (`AddTrials`)."""

  def __init__(self, study_config):
    self.study_config = study_config
    self.trials: List[Any] = []

  def AddTrials(self, trials: Sequence[Any]) -> None:   # pylint: disable=invalid-name
    for t in trials:
      if not getattr(t, 'id', 0):
        t.id = len(self.trials) + 1
      self.trials.append(t)

  def GetTrials(self, *, status_matches=None, **_unused) -> List[Any]:   # pylint: disable=invalid-name
    if status_matches is None:
      return list(self.trials)
    want = getattr(status_matches, 'name', str(status_matches))
    out = []
    for t in self.trials:
      status = 'COMPLETED' if t.is_completed else 'ACTIVE'
      if status == want:
        out.append(t)
    return out


class DesignerPolicy:
  """designer_policy.py:44-125."""

  def __init__(self, supporter, designer_factory: Callable[[Any], Any], *, policy_name: str = 'DesignerPolicy'):
    self._supporter = supporter
    self._designer_factory = designer_factory
    self._policy_name = policy_name
    self._designer = None

  def suggest(self, request: SuggestRequest) -> SuggestDecision:
    designer = self._designer_factory(request.study_config)
    completed = self._supporter.GetTrials(status_matches=vz.TrialStatus.COMPLETED)
    active = self._supporter.GetTrials(status_matches=vz.TrialStatus.ACTIVE)
    designer.update(vz.CompletedTrials(completed), vz.ActiveTrials(active))
    self._designer = designer   # saved for debugging purposes only (designer_policy.py:104)
    return SuggestDecision(designer.suggest(request.count))

  def early_stop(self, request):
    raise NotImplementedError('DesignerPolicy does not support the early_stop() method.')

  @property
  def name(self) -> str:
    return self._policy_name
