__version__ = '0.1.1-b200'
