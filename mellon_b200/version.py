"""Version information (API level of the reference this package mirrors: mellon 1.7.1)."""

__version__ = "1.7.1+b200.1"
