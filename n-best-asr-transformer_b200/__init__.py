"""B200-native (sm_100a) hot path of N-Best-ASR-Transformer behind the reference's Python API.

Submodules (import as `nbest_b200.<name>`):
  _lib      ctypes binding of libnbest_sm100.so (include/nbest_sm100.h)
  ops       thin torch-tensor wrappers over the C ABI (device pointers + current stream)
  model     make_model / TOD_ASR_Transformer_STC drop-in (reference models/model.py)
  optim     BertAdam drop-in (reference models/optimization.py)
  inputs    prepare_inputs_for_roberta drop-in + packed layout (reference utils/bert_xlnet_inputs.py)
  trainer   8-rank data-parallel trainer (replaces reference utils/gpu_selection.py)
"""
__version__ = "0.1.0"
