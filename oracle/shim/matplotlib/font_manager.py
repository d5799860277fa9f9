''' Stand-in for matplotlib.font_manager -- TEST INFRASTRUCTURE ONLY. '''
class _FM:
    ttflist = []
    def addfont(self, *a, **k): pass
fontManager = _FM()
def findfont(*a, **k):
    raise ValueError('no fonts in the shim')
def findSystemFonts(*a, **k):
    return []
