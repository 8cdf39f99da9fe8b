"""A small, dependency-free stand-in for the slice of `vizier.pyvizier` the GP-bandit designer uses.

The drop-in target is the real package: when `vizier.pyvizier` is importable, `vizier_b200.vz`
re-exports it and this module is unused.  It exists because the reference's `vizier.pyvizier`
needs generated protobuf modules that cannot be built offline, and the GPU test box
has no reference checkout at all.  Names, argument meaning and error behaviour follow
vizier/_src/pyvizier/shared/{parameter_config,base_study_config,trial,common}.py; the
implementation is new and deliberately minimal (flat search spaces, single-process use).
"""

from __future__ import annotations

import copy
import dataclasses
import datetime
import abc
import enum
import math
from typing import Any, Dict, Iterable, Iterator, List, Mapping, MutableMapping, Optional, Sequence, Tuple, Union


class ParameterType(enum.Enum):
  DOUBLE = 'DOUBLE'
  INTEGER = 'INTEGER'
  CATEGORICAL = 'CATEGORICAL'
  DISCRETE = 'DISCRETE'
  CUSTOM = 'CUSTOM'

  def is_numeric(self) -> bool:
    return self in (ParameterType.DOUBLE, ParameterType.INTEGER, ParameterType.DISCRETE)

  def is_continuous(self) -> bool:
    return self == ParameterType.DOUBLE


class ScaleType(enum.Enum):
  LINEAR = 'LINEAR'
  LOG = 'LOG'
  REVERSE_LOG = 'REVERSE_LOG'
  UNIFORM_DISCRETE = 'UNIFORM_DISCRETE'


class ObjectiveMetricGoal(enum.IntEnum):
  MAXIMIZE = 1
  MINIMIZE = 2

  @property
  def is_maximize(self) -> bool:
    return self == ObjectiveMetricGoal.MAXIMIZE

  @property
  def is_minimize(self) -> bool:
    return self == ObjectiveMetricGoal.MINIMIZE


ParameterValueTypes = Union[str, int, float, bool]


@dataclasses.dataclass(frozen=True)
class ParameterValue:
  value: ParameterValueTypes

  def cast_as_internal(self, internal_type: ParameterType) -> ParameterValueTypes:
    if internal_type == ParameterType.DOUBLE or internal_type == ParameterType.DISCRETE:
      return float(self.value)
    if internal_type == ParameterType.INTEGER:
      return int(self.value)
    if internal_type == ParameterType.CATEGORICAL:
      return str(self.value)
    return self.value

  @property
  def as_float(self) -> Optional[float]:
    return float(self.value) if isinstance(self.value, (int, float)) and not isinstance(self.value, bool) else None

  @property
  def as_str(self) -> Optional[str]:
    return self.value if isinstance(self.value, str) else None


class ParameterDict(MutableMapping):
  """name -> ParameterValue; raw python values are wrapped on assignment."""

  def __init__(self, iterable: Any = (), **kwargs):
    self._items: Dict[str, ParameterValue] = {}
    self.update(iterable, **kwargs)

  def __setitem__(self, key: str, value: Union[ParameterValue, ParameterValueTypes]):
    self._items[key] = value if isinstance(value, ParameterValue) else ParameterValue(value)

  def __getitem__(self, key: str) -> ParameterValue:
    return self._items[key]

  def __delitem__(self, key: str):
    del self._items[key]

  def __iter__(self) -> Iterator[str]:
    return iter(self._items)

  def __len__(self) -> int:
    return len(self._items)

  def get_value(self, key: str, default: Any = None) -> Any:
    pv = self._items.get(key)
    return default if pv is None else pv.value

  def as_dict(self) -> Dict[str, ParameterValueTypes]:
    return {k: v.value for k, v in self._items.items()}

  def __repr__(self) -> str:
    return f'ParameterDict({self.as_dict()!r})'


@dataclasses.dataclass
class ParameterConfig:
  """Flat (non-conditional) parameter description."""

  name: str
  type: ParameterType
  bounds: Optional[Tuple[float, float]] = None
  feasible_values: Sequence[Any] = ()
  scale_type: Optional[ScaleType] = None
  default_value: Optional[Any] = None

  @classmethod
  def factory(cls, name: str, *, bounds=None, feasible_values=None, scale_type=None, default_value=None) -> 'ParameterConfig':
    if not name:
      raise ValueError('Parameter name cannot be empty.')
    if (bounds is None) == (feasible_values is None):
      raise ValueError('Exactly one of "bounds" or "feasible_values" must be provided.')
    if bounds is not None:
      lo, hi = bounds
      if lo > hi:
        raise ValueError(f'Lower bound {lo} exceeds upper bound {hi} for {name}.')
      if isinstance(lo, int) and isinstance(hi, int) and not isinstance(lo, bool):
        return cls(name, ParameterType.INTEGER, (int(lo), int(hi)), tuple(range(int(lo), int(hi) + 1)), scale_type, default_value)
      return cls(name, ParameterType.DOUBLE, (float(lo), float(hi)), (), scale_type, default_value)
    vals = list(feasible_values)
    if not vals:
      raise ValueError('feasible_values cannot be empty.')
    if all(isinstance(v, str) for v in vals):
      return cls(name, ParameterType.CATEGORICAL, None, tuple(sorted(vals)), scale_type, default_value)
    vals = sorted(float(v) for v in vals)
    return cls(name, ParameterType.DISCRETE, (vals[0], vals[-1]), tuple(vals), scale_type, default_value)

  @property
  def num_feasible_values(self) -> Union[int, float]:
    if self.type == ParameterType.DOUBLE:
      return float('inf')
    return len(self.feasible_values)

  def continuify(self) -> 'ParameterConfig':
    if self.type == ParameterType.DOUBLE:
      return copy.deepcopy(self)
    if not self.type.is_numeric():
      raise ValueError(f'Cannot continuify {self.type}')
    scale = self.scale_type
    if scale == ScaleType.UNIFORM_DISCRETE:
      scale = None
    return ParameterConfig(self.name, ParameterType.DOUBLE, (float(self.bounds[0]), float(self.bounds[1])), (), scale, self.default_value)

  def contains(self, value: Union[ParameterValue, ParameterValueTypes]) -> bool:
    v = value.value if isinstance(value, ParameterValue) else value
    if self.type == ParameterType.DOUBLE:
      return isinstance(v, (int, float)) and self.bounds[0] <= v <= self.bounds[1]
    return v in self.feasible_values


class _Root:
  """`search_space.root` / `select_root()` selector with the add_*_param helpers."""

  def __init__(self, space: 'SearchSpace'):
    self._space = space

  def _add(self, pc: ParameterConfig) -> ParameterConfig:
    if pc.name in self._space._by_name:
      raise ValueError(f'Duplicate parameter name: {pc.name}')
    self._space._by_name[pc.name] = pc
    return pc

  def add_float_param(self, name: str, min_value: float, max_value: float, *, default_value=None,
                      scale_type: Optional[ScaleType] = ScaleType.LINEAR, index=None) -> ParameterConfig:
    del index
    return self._add(ParameterConfig.factory(name, bounds=(float(min_value), float(max_value)), scale_type=scale_type, default_value=default_value))

  def add_int_param(self, name: str, min_value: int, max_value: int, *, default_value=None,
                    scale_type: Optional[ScaleType] = None, index=None) -> ParameterConfig:
    del index
    if int(min_value) != min_value or int(max_value) != max_value:
      raise ValueError('min_value and max_value must be integers.')
    return self._add(ParameterConfig.factory(name, bounds=(int(min_value), int(max_value)), scale_type=scale_type, default_value=default_value))

  def add_discrete_param(self, name: str, feasible_values: Sequence[float], *, default_value=None,
                         scale_type: Optional[ScaleType] = ScaleType.LINEAR, index=None, auto_cast=True) -> ParameterConfig:
    del index, auto_cast
    return self._add(ParameterConfig.factory(name, feasible_values=[float(v) for v in feasible_values], scale_type=scale_type, default_value=default_value))

  def add_categorical_param(self, name: str, feasible_values: Sequence[str], *, default_value=None, scale_type=None, index=None) -> ParameterConfig:
    del index, scale_type
    return self._add(ParameterConfig.factory(name, feasible_values=[str(v) for v in feasible_values], default_value=default_value))

  def add_bool_param(self, name: str, feasible_values=None, *, default_value=None, scale_type=None, index=None) -> ParameterConfig:
    del feasible_values, index, scale_type
    return self._add(ParameterConfig.factory(name, feasible_values=['False', 'True'], default_value=default_value))


class SearchSpace:
  def __init__(self):
    self._by_name: Dict[str, ParameterConfig] = {}

  @property
  def root(self) -> _Root:
    return _Root(self)

  def select_root(self) -> _Root:
    return _Root(self)

  @property
  def parameters(self) -> List[ParameterConfig]:
    return list(self._by_name.values())

  def get(self, name: str) -> ParameterConfig:
    return self._by_name[name]

  def num_parameters(self, param_type: Optional[ParameterType] = None) -> int:
    if param_type is None:
      return len(self._by_name)
    return sum(1 for p in self._by_name.values() if p.type == param_type)

  @property
  def is_conditional(self) -> bool:
    return False

  def contains(self, parameters: Mapping[str, Any]) -> bool:
    try:
      return set(parameters.keys()) == set(self._by_name) and all(
          self._by_name[k].contains(v) for k, v in parameters.items())
    except KeyError:
      return False


@dataclasses.dataclass
class MetricInformation:
  name: str = ''
  goal: ObjectiveMetricGoal = ObjectiveMetricGoal.MAXIMIZE
  safety_threshold: Optional[float] = None
  min_value: float = -math.inf
  max_value: float = math.inf

  def flip_goal(self) -> 'MetricInformation':
    g = ObjectiveMetricGoal.MINIMIZE if self.goal.is_maximize else ObjectiveMetricGoal.MAXIMIZE
    return dataclasses.replace(self, goal=g)


class MetricsConfig(list):
  def item(self) -> MetricInformation:
    if len(self) != 1:
      raise ValueError(f'Found {len(self)} metrics; item() needs exactly one.')
    return self[0]

  @property
  def is_single_objective(self) -> bool:
    return len([m for m in self if m.safety_threshold is None]) == 1

  def of_type(self, *_args, **_kwargs) -> 'MetricsConfig':
    return MetricsConfig(self)


class Metadata(MutableMapping):
  """Namespaced string key-value store: `md.ns('a').ns('b')['k'] = 'v'`."""

  def __init__(self, *args, **kwargs):
    self._stores: Dict[Tuple[str, ...], Dict[str, Any]] = {(): {}}
    self._ns: Tuple[str, ...] = ()
    self.update(dict(*args, **kwargs))

  def ns(self, component: str) -> 'Metadata':
    child = Metadata.__new__(Metadata)
    child._stores = self._stores
    child._ns = self._ns + (component,)
    child._stores.setdefault(child._ns, {})
    return child

  def abs_ns(self, namespace: Iterable[str] = ()) -> 'Metadata':
    child = Metadata.__new__(Metadata)
    child._stores = self._stores
    child._ns = tuple(namespace)
    child._stores.setdefault(child._ns, {})
    return child

  def namespaces(self) -> List[Tuple[str, ...]]:
    return [k for k, v in self._stores.items() if v]

  def __getitem__(self, key: str):
    return self._stores[self._ns][key]

  def __setitem__(self, key: str, value):
    self._stores.setdefault(self._ns, {})[key] = value

  def __delitem__(self, key: str):
    del self._stores[self._ns][key]

  def __iter__(self):
    return iter(self._stores.get(self._ns, {}))

  def __len__(self):
    return len(self._stores.get(self._ns, {}))

  def __repr__(self):
    return f'Metadata({ {"/".join(k): v for k, v in self._stores.items() if v} })'


@dataclasses.dataclass
class ProblemStatement:
  search_space: SearchSpace = dataclasses.field(default_factory=SearchSpace)
  metric_information: MetricsConfig = dataclasses.field(default_factory=MetricsConfig)
  metadata: Metadata = dataclasses.field(default_factory=Metadata)

  @property
  def is_single_objective(self) -> bool:
    return self.metric_information.is_single_objective


@dataclasses.dataclass
class Metric:
  value: float
  std: Optional[float] = None


class Measurement:
  def __init__(self, metrics: Optional[Mapping[str, Union[float, Metric]]] = None, elapsed_secs: float = 0.0, steps: float = 0.0):
    self.metrics: Dict[str, Metric] = {}
    for k, v in (metrics or {}).items():
      self.metrics[k] = v if isinstance(v, Metric) else Metric(float(v))
    self.elapsed_secs = elapsed_secs
    self.steps = steps


class TrialSuggestion:
  def __init__(self, parameters: Any = None, *, metadata: Optional[Metadata] = None):
    self.parameters = parameters if isinstance(parameters, ParameterDict) else ParameterDict(parameters or {})
    self.metadata = metadata if metadata is not None else Metadata()

  def to_trial(self, uid: int = 0) -> 'Trial':
    return Trial(id=uid, parameters=self.parameters, metadata=self.metadata)


_counter = [0]


class Trial(TrialSuggestion):
  def __init__(self, parameters: Any = None, *, id: int = 0, metadata: Optional[Metadata] = None,  # pylint: disable=redefined-builtin
               final_measurement: Optional[Measurement] = None, infeasibility_reason: Optional[str] = None,
               creation_time: Optional[datetime.datetime] = None):
    super().__init__(parameters, metadata=metadata)
    self.id = id
    self.final_measurement = final_measurement
    self.infeasibility_reason = infeasibility_reason
    self.measurements: List[Measurement] = []
    # creation order breaks ties when timestamps coincide
    _counter[0] += 1
    self._seq = _counter[0]
    self.creation_time = creation_time or datetime.datetime.now()
    self.completion_time: Optional[datetime.datetime] = None

  @property
  def infeasible(self) -> bool:
    return self.infeasibility_reason is not None

  @property
  def is_completed(self) -> bool:
    return self.final_measurement is not None or self.infeasible

  def complete(self, measurement: Optional[Measurement] = None, *, infeasibility_reason: Optional[str] = None, inplace: bool = True) -> 'Trial':
    t = self if inplace else copy.deepcopy(self)
    t.final_measurement = measurement
    t.infeasibility_reason = infeasibility_reason
    t.completion_time = datetime.datetime.now()
    return t


@dataclasses.dataclass(frozen=True)
class CompletedTrials:
  """vizier/_src/algorithms/core/abstractions.py:32-50."""

  trials: Sequence[Trial] = ()

  def __post_init__(self):
    object.__setattr__(self, 'trials', tuple(self.trials))
    for t in self.trials:
      if not t.is_completed:
        raise ValueError(f'All trials must be completed. Bad trial: {t}')


@dataclasses.dataclass(frozen=True)
class ActiveTrials:
  """vizier/_src/algorithms/core/abstractions.py:53-71."""

  trials: Sequence[Trial] = ()

  def __post_init__(self):
    object.__setattr__(self, 'trials', tuple(self.trials))
    for t in self.trials:
      if t.is_completed:
        raise ValueError(f'All trials must be active. Bad trial: {t}')


@dataclasses.dataclass
class Prediction:
  """vizier/_src/algorithms/core/abstractions.py:152-171."""

  mean: Any
  stddev: Any
  metadata: Optional[Metadata] = None

  def __post_init__(self):
    if getattr(self.mean, 'shape', None) != getattr(self.stddev, 'shape', None):
      raise ValueError('mean and stddev must have the same shape')


class Designer(abc.ABC):
  """vizier/_src/algorithms/core/abstractions.py:74-149 (`_SuggestionAlgorithm` + `Designer`): the
  suggest/update contract `DesignerPolicy` drives."""

  @abc.abstractmethod
  def suggest(self, count: Optional[int] = None) -> Sequence[TrialSuggestion]:
    """Makes `count` suggestions (None: as many as the algorithm wants)."""

  @abc.abstractmethod
  def update(self, completed: CompletedTrials, all_active: ActiveTrials) -> None:
    """Incorporates newly COMPLETED trials and the full list of ACTIVE trials."""


class Predictor(abc.ABC):
  """vizier/_src/algorithms/core/abstractions.py:174-199."""

  @abc.abstractmethod
  def predict(self, trials: Sequence[TrialSuggestion], rng: Any = None, num_samples: Optional[int] = None) -> Prediction:
    """Mean and stddev at the given suggestions."""


class TrialStatus(enum.Enum):
  """vizier/_src/pyvizier/shared/trial.py: the statuses DesignerPolicy filters by."""
  UNKNOWN = 'UNKNOWN'
  REQUESTED = 'REQUESTED'
  ACTIVE = 'ACTIVE'
  COMPLETED = 'COMPLETED'
  STOPPING = 'STOPPING'
