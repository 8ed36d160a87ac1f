from .data_generator import (ArraySource, HDF5Source, PrefetchSessionsGenerator, ConcatSessionsGenerator,  # noqa: F401
                             split_trials)
