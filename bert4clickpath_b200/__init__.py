"""bert4clickpath_b200 — B200-native hot path of the clickstream transformer
(MiladShahidi/BERT4ClickPath), behind the reference's Python construction surface.

Importing the package does not need a GPU; constructing a model does (there is no CPU path).
"""
from .constants import *  # noqa: F401,F403
from .constants import LABEL_PAD, INPUT_MASKING_TOKEN, RESERVED_TOKENS  # noqa: F401


def __getattr__(name):
    # heavy modules (torch + CUDA) are imported lazily
    import importlib
    table = {
        "ClickstreamTransformer": ".clickstream_transformer",
        "TransformerInputPrep": ".clickstream_transformer",
        "Adam": ".clickstream_transformer",
        "load_vocabulary": ".clickstream_transformer",
        "Transformer": ".transformer",
        "positional_encoding": ".transformer",
        "create_padding_mask": ".transformer",
        "create_segment_markers": ".transformer",
        "SoftMaxHead": ".head",
        "BinaryClassificationHead": ".head",
        "MultiLabel_MultiClass_classification": ".head",
        "MaskedLoss": ".losses",
        "sparse_categorical_crossentropy": ".losses",
        "binary_crossentropy": ".losses",
        "ClozeMaskedLoss": ".cloze",
        "ClozeMaskedRecall": ".cloze",
        "ClozeMaskedNDCG": ".cloze",
        "cloze_output_adaptor": ".cloze",
        "PositiveRate": ".metrics",
        "PredictedPositives": ".metrics",
        "F1Score": ".metrics",
        "MaskedMetric": ".metrics",
        "create_cloze_dataset": ".data",
        "ClozeDataset": ".data",
        "DeviceClozeBuilder": ".data",
        "CustomLRSchedule": ".training_utils",
        "CustomExponentialDecayLR": ".training_utils",
        "BestModelSaverCallback": ".training_utils",
        "ReduceLROnPlateau": ".training_utils",
        "EarlyStopping": ".training_utils",
    }
    if name in table:
        return getattr(importlib.import_module(table[name], __name__), name)
    raise AttributeError(name)
