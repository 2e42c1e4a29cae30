# Mirrors seekr/__version__.py:4 of the reference this package drops in for.
__version__ = "2.0.2"
__b200_version__ = "0.1.0"
