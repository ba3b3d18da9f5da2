"""gnina_b200 — B200-native CNN-scoring hot path of gnina behind gnina's own scorer interface.

Product code: CUDA kernels + C ABI in csrc/ (libgnina_b200.so), host mirror of the reference interface in
scorer.py.  Nothing here imports the CPU oracle (oracle/ is test infrastructure)."""
from .scorer import CNNScorer, usage_error, expand_model_names, builtin_models, PRECISION_FP32, PRECISION_FP16_TC  # noqa: F401
from .batching import PoseQueue  # noqa: F401,E402
from .gninatypes import read_gninatypes, write_gninatypes  # noqa: F401,E402
