'''Shared list of parity cases: (parameters, grid, state).  Used by make_golden.py (run once
in the build container against the unmodified Python reference) and by the tests (which load
the committed .npz files; nothing here touches /root/reference).'''
import numpy

LDC = {'Reynolds Number': 100}
RB = {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 1000.0, 'Prandtl Number': 10.0,
      'Biot Number': 1.0, 'X-max': 10, 'Y-max': 10, 'Reynolds Number': 1}
RBP = {'Problem Type': 'Rayleigh-Benard Perturbation', 'Rayleigh Number': 1500.0, 'Prandtl Number': 10.0,
       'Biot Number': 1.0, 'X-max': 10, 'Asymmetry Parameter': 0.3}
DHC = {'Problem Type': 'Differentially Heated Cavity', 'Rayleigh Number': 1e4, 'Prandtl Number': 1000.0,
       'Reynolds Number': 1, 'X-max': 0.051, 'Y-max': 1}
QG = {'Problem Type': 'Double Gyre', 'Reynolds Number': 16, 'Rossby Parameter': 1000,
      'Wind Stress Parameter': 1000}
AMOC = {'Problem Type': 'AMOC', 'Rayleigh Number': 4e4, 'Prandtl Number': 2.25, 'Lewis Number': 1,
        'Freshwater Flux': 0.1, 'Temperature Forcing': 1, 'X-max': 5}
STR = {'Grid Stretching Factor': 1.5}

# name: (parameters, nx, ny, nz, dim, dof, state-kind)   state-kind: 'lin' | 'zero' | int seed
CASES = {
    'ldc3d_lin': (LDC, 4, 4, 4, None, None, 'lin'),
    'ldc3d_rand': (LDC, 6, 5, 4, None, None, 0),
    'ldc3d_zero': (LDC, 4, 4, 4, None, None, 'zero'),
    'ldc3d_str_rand': ({**LDC, **STR, 'Lid Velocity': 2.5}, 5, 6, 7, None, None, 1),
    'ldc3d_8_rand': (LDC, 8, 8, 8, None, None, 2),
    'ldc3d_small': (LDC, 2, 3, 2, None, None, 3),
    'stokes3d_rand': ({'Reynolds Number': 0}, 4, 4, 4, None, None, 4),
    'ldc2d_rand': ({**LDC, **STR}, 8, 6, 1, None, None, 5),
    'ldc2d_lin': (LDC, 4, 4, 1, None, None, 'lin'),
    'ldc_semi2d_rand': (LDC, 6, 5, 1, 3, 4, 6),
    'rb3d_rand': (RB, 4, 5, 6, None, None, 7),
    'rb3d_str_lin': ({**RB, **STR}, 4, 4, 4, None, None, 'lin'),
    'rb2d_rand': (RB, 8, 6, 1, None, None, 8),
    'rbp3d_rand': (RBP, 5, 4, 6, None, None, 9),
    'rbp2d_str_rand': ({**RBP, **STR}, 6, 8, 1, None, None, 10),
    'rb_semi2d_rand': (RB, 6, 5, 1, 3, 5, 11),
    'dhc2d_rand': (DHC, 8, 8, 1, None, None, 12),
    'dhc3d_rand': (DHC, 4, 5, 4, None, None, 13),
    'dhc_semi2d_rand': (DHC, 6, 5, 1, 3, 5, 16),
    'qg_rand': (QG, 8, 6, 1, None, None, 14),
    'qg_zero': (QG, 6, 6, 1, None, None, 'zero'),
    'amoc_rand': (AMOC, 8, 6, 1, None, None, 15),
    'amoc_str_lin': ({**AMOC, **STR}, 6, 4, 1, None, None, 'lin'),
}


def make_state(kind, n):
    if kind == 'lin':
        return numpy.arange(1, n + 1, dtype=numpy.float64)
    if kind == 'zero':
        return numpy.zeros(n)
    return numpy.random.default_rng(kind).uniform(-0.5, 0.5, n)


# ---- user-supplied boundary conditions (Discretization.py:62-66; reference test: tests/test_interface.py:117-150) ----
# name: (parameters, nx, ny, nz, dim, dof, state-kind, callback).  The callbacks only use the public methods of the
# reference's BoundaryConditions class; the same functions drive the reference (make_golden.py) and the B200 Interface.
def _bc_reference_test(bc, atom):
    '''The callback of the reference's own test_custom_bc.'''
    bc.heat_flux_east(atom, 0)
    bc.heat_flux_west(atom, 0)
    bc.no_slip_east(atom)
    bc.no_slip_west(atom)
    bc.heat_flux_north(atom, 0)
    bc.heat_flux_south(atom, 0)
    bc.no_slip_north(atom)
    bc.no_slip_south(atom)
    Bi = 0
    bc.heat_flux_top(atom, 0, Bi)
    bc.temperature_bottom(atom, 0)
    bc.free_slip_top(atom)
    bc.no_slip_bottom(atom)
    return bc.get_forcing()


def _bc_heated_box(bc, atom):
    '''Same order of ops, every constant non-trivial: side-wall heat fluxes, a Robin lid, a hot bottom.'''
    bc.heat_flux_east(atom, 0.25)
    bc.heat_flux_west(atom, -0.5, 0.2)
    bc.no_slip_east(atom)
    bc.no_slip_west(atom)
    bc.heat_flux_north(atom, 0.1, 0.3)
    bc.heat_flux_south(atom, 0)
    bc.no_slip_north(atom)
    bc.no_slip_south(atom)
    bc.heat_flux_top(atom, 0.4, 1.5)
    bc.temperature_bottom(atom, 2.0)
    bc.free_slip_top(atom)
    bc.no_slip_bottom(atom)
    return bc.get_forcing()


def _bc_fast_lid(bc, atom):
    '''Lid-driven cavity with the lid velocity given by the callback instead of the parameter list.'''
    bc.no_slip_east(atom)
    bc.no_slip_west(atom)
    bc.no_slip_south(atom)
    bc.no_slip_north(atom)
    bc.no_slip_bottom(atom)
    bc.moving_lid_top(atom, 3.5)
    return bc.get_forcing()


def _bc_walls_2d(bc, atom):
    '''2-D box with wall temperatures of the callback's choosing (order of ops of the heated cavity).'''
    bc.temperature_east(atom, -0.75)
    bc.temperature_west(atom, 1.25)
    bc.no_slip_east(atom)
    bc.no_slip_west(atom)
    bc.heat_flux_north(atom, 0.2)
    bc.heat_flux_south(atom, 0, 0.5)
    bc.no_slip_north(atom)
    bc.no_slip_south(atom)
    return bc.get_forcing()


def _bc_unsupported(bc, atom):
    '''An order of ops no kernel family is generated for (free-slip side walls under a moving lid).'''
    bc.free_slip_east(atom)
    bc.free_slip_west(atom)
    bc.no_slip_south(atom)
    bc.moving_lid_north(atom, 1.0)
    return bc.get_forcing()


CUSTOM_BC_CASES = {
    'custom_reference_test': ({'Rayleigh Number': 100, 'Prandtl Number': 100}, 4, 4, 4, 3, 5, 20, _bc_reference_test),
    'custom_heated_box': ({'Rayleigh Number': 800.0, 'Prandtl Number': 7.0, 'Reynolds Number': 1, **STR}, 5, 4, 6, 3, 5, 21,
                          _bc_heated_box),
    'custom_fast_lid': ({'Reynolds Number': 40}, 5, 6, 4, 3, 4, 22, _bc_fast_lid),
    'custom_walls_2d': ({'Rayleigh Number': 1e3, 'Prandtl Number': 10.0, 'Reynolds Number': 1}, 7, 6, 1, 2, 4, 23,
                        _bc_walls_2d),
}
