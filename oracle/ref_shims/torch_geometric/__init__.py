"""Stand-in for torch_geometric==1.6.3 (only the symbols the reference hot path imports)."""
