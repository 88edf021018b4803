"""Top-level `models` module: the drop-in for the reference's models.py.

hparams/pretrain.yaml:84 / finetune.yaml:85 instantiate `!new:models.ScoreTranscription`; pretrain.py / finetune.py do
`from models import ...`-style access to `labels`, `SOS`, `EOS`, `vocab_size`.  Put this repository first on
`sys.path` and those scripts run on the B200-native implementation unchanged.
"""
from piano_a2s_b200.models import (  # noqa: F401
    AttentionLayer, ConvStack, Encoder, HierarchicalDecoder, LabelsMultiple, NoteDecoder, ScoreTranscription,
    EOS, PAD, SOS, init_bn, init_gru, init_layer, labels, vocab_size,
)
