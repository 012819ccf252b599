"""Drop-in for the reference's ``models`` package (same import paths)."""
