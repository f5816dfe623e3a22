"""Empty stand-in: imageio is only used by the reference's image dumps / metrics / mesh export."""
