'''Newton / linear-solve parity cases (reference: SciPy backend, SuperLU).'''
NEWTON_CASES = {
    'ldc3d_8': ({'Reynolds Number': 100, 'Lid Velocity': 1}, 8, 8, 8),
    'ldc3d_12_str': ({'Reynolds Number': 50, 'Lid Velocity': 1, 'Grid Stretching Factor': 1.5}, 12, 10, 8),
    'ldc2d_24': ({'Reynolds Number': 200, 'Lid Velocity': 1}, 24, 24, 1),
    'dhc2d_16': ({'Problem Type': 'Differentially Heated Cavity', 'Rayleigh Number': 1e4, 'Prandtl Number': 1000.0,
                  'Reynolds Number': 1, 'X-max': 0.051, 'Y-max': 1}, 16, 16, 1),
    'rb3d_8': ({'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 500.0, 'Prandtl Number': 10.0,
                'Biot Number': 1.0, 'X-max': 10, 'Y-max': 10}, 8, 8, 6),
    'amoc_16': ({'Problem Type': 'AMOC', 'Rayleigh Number': 4e4, 'Prandtl Number': 2.25, 'Lewis Number': 1,
                 'Freshwater Flux': 0.0, 'Temperature Forcing': 1, 'X-max': 5}, 16, 8, 1),
    'qg_16': ({'Problem Type': 'Double Gyre', 'Reynolds Number': 16, 'Rossby Parameter': 100,
               'Wind Stress Parameter': 100}, 16, 16, 1),
}
