"""pycuda.autoinit: nothing to initialise (torch owns the CUDA context)."""
