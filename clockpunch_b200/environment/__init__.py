"""Environment layer on the CUDA step: SSAScheduler mirror and its helpers."""
