"""Empty stand-in: plyfile is only used by the reference's image dumps / metrics / mesh export."""
