"""pepr_b200 -- B200-native maximum-likelihood engine behind PEPR's RAxMLRunner / FastTreeRunner interface.

Host code above the C ABI (include/peprml.h).  The compute path is the CUDA library pepr_b200/libpeprml.so; there is
no CPU fallback: importing works without a GPU (so host logic can be tested), creating a Context does not.
"""
from .engine import (Alignment, Context, EngineError, Group, Tree, bootstrap_weights, constraints_from_tree, crunch_patterns, gamma_rates,  # noqa: F401
                     newick_satisfies_constraints, parsimony_tree, parsimony_tree_constrained, pattern_range, lib, support_counts,
                     support_tree, unique_id, wag_frequencies, wag_pmatrix)

__all__ = ["constraints_from_tree", "newick_satisfies_constraints", "parsimony_tree_constrained", "Alignment", "Context", "EngineError", "Group", "Tree", "bootstrap_weights", "crunch_patterns", "pattern_range", "parsimony_tree", "gamma_rates", "lib", "support_counts",
           "support_tree", "unique_id", "wag_frequencies", "wag_pmatrix"]
