"""Drop-in shadows of the reference's `models` package: each module re-exports tensoir_b200 under the reference's names (option A)."""
