"""Rule-author API under the reference's module names (legacy ``vescale/dtensor/ops/``; new ``vescale/dtensor/_ops/``).

The framework's own rules live in ``dtensor/rules/`` and speak ``RuleResult``.  This package is for code written against the
reference's contract — ``register_prop_rule`` (rule: schema -> ``OutputSharding`` with optional ``schema_suggestions``) and
``register_op_strategy`` (strategy: (mesh, schema of ``OpStrategy`` arguments) -> ``OpStrategy`` of alternatives with redistribute
costs) — plus the building blocks such rules are made of: the einsum-notation propagation ``einop_rule`` / ``pointwise_rule``
(``common_rules.py``), the einsum strategy generator (``basic_strategy.py``) and the spec predicates (``utils.py``).  Rules registered
here land in the same registry as the built-in ones and take precedence over them."""
from . import basic_strategy, common_rules, utils  # noqa: F401
from .utils import register_op_strategy, register_prop_rule  # noqa: F401
