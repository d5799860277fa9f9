''' Stand-in for matplotlib -- TEST INFRASTRUCTURE ONLY. '''
__version__ = '0.0'
