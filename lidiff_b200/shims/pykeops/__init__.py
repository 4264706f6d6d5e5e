__version__ = "2.1.2+lidiff_b200"
