"""smftools_b200 -- B200-native (sm_100a) implementation of the smftools HMM footprinting hot path.

Public surface mirrors ``smftools.hmm.HMM`` of the reference (see ``smftools_b200/hmm.py``).
"""
from .hmm import (  # noqa: F401
    HMM,
    BaseHMM,
    CodedU8,
    DistanceBinnedSingleBernoulliHMM,
    MultiBernoulliHMM,
    SingleBernoulliHMM,
    create_hmm,
    install_into_reference,
    mask_layers_outside_read_span,
    normalize_hmm_feature_sets,
    register_hmm,
)

from .merges import apply_merges, expand_merge_chain, expand_merge_chain_device, merge_chain_packed  # noqa: F401,E402

from .inputs import (  # noqa: F401,E402
    build_multi_channel_union,
    build_single_channel,
    mask_uncovered_model_input,
    prepare_model_input,
    resolve_pos_mask_for_methbase,
)

__version__ = "0.2.0"
