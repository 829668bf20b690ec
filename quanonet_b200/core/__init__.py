"""Host-side mirror of the reference's ``core`` package for the PyTorch quantum path."""
