from .linear import (DampedSkewSymmetricLinear, DerivedWeight, IdentityInitLinear, IdentityLinear, LassoLinearRELU,  # noqa: F401
                     LeftStochasticLinear, Linear, LinearBase, NonNegativeLinear, PerronFrobeniusLinear, PSDLinear,
                     RightStochasticLinear, SkewSymmetricLinear, SplitLinear, SquareLinear, StableSplitLinear, SVDLinear,
                     SVDLinearLearnBounds, SymmetricLinear, SymmetricSVDLinear, maps, square_maps)
