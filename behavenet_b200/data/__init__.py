from .data_generator import (ArraySource, HDF5Source, PrefetchSessionsGenerator, PrefetchSessionsGeneratorMulti,  # noqa: F401
                             ConcatSessionsGenerator, ConcatSessionsGeneratorMulti, split_trials)
