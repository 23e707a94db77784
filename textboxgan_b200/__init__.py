"""B200-native TextBoxGAN training-step hot path."""
import torch as _torch

# The few plain library GEMMs left on the path (LSTM input projections and attention keys of the frozen recogniser,
# the second-order dense nodes of the R1 pass) run on TF32 tensor cores; the mapping network, style projections and
# discriminator dense layers are exact fp32 on this repo's kernels, everything heavy is bf16 on its tensor-core kernels.
_torch.backends.cuda.matmul.allow_tf32 = True
