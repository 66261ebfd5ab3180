__version__ = '0.1.0+b200'
abi_version = 1
