"""R14-R16: SD / ControlNet / VAE-encoder guidance on tcgen05 kernels (see model.py, guidance.py)."""
