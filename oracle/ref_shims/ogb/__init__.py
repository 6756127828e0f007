"""Stand-in for ogb==1.3.1 (AtomEncoder/BondEncoder only)."""
