"""Stand-in."""
