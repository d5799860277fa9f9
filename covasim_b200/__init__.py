'''
covasim_b200 -- B200-native implementation of Covasim's per-timestep simulation hot path.
(Package body is filled in as the build proceeds; see DESIGN.md.)
'''
__version__ = '0.1.0'
