"""shade / eq_hist on the GPU (transfer_functions/__init__.py of the reference) - filled in below."""
