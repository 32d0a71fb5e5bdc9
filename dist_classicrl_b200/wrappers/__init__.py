from dist_classicrl_b200.wrappers.dummy_vec_wrapper import DummyVecWrapper
from dist_classicrl_b200.wrappers.flatten_multidiscrete_wrapper import (FlattenMultiDiscreteActionsWrapper,
                                                                       FlattenMultiDiscreteObservationsWrapper)

__all__ = ["DummyVecWrapper", "FlattenMultiDiscreteActionsWrapper", "FlattenMultiDiscreteObservationsWrapper"]
