"""B200-native TextBoxGAN training-step hot path."""
import torch as _torch

# The few plain library GEMMs left on the path (mapping MLP, mod_dense, discriminator dense layers,
# LSTM input projections) run on TF32 tensor cores; everything heavy is bf16 on this repo's kernels.
_torch.backends.cuda.matmul.allow_tf32 = True
