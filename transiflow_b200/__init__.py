'''transiflow_b200 -- B200-native computational backend for TransiFlow's interface API.'''
from .interface import DeviceMatrix, Interface  # noqa: F401


def create(parameters, nx, ny, nz=1, dim=None, dof=None, x=None, y=None, z=None,
           boundary_conditions=None, backend='B200', **kwargs):
    '''Drop-in for ``transiflow.interface.create`` (interface/create.py:4-68) with one more
    branch: ``backend='B200'``.  Other backend names are forwarded to the reference.'''
    if backend.lower() == 'b200':
        return Interface(parameters, nx, ny, nz, dim, dof, x, y, z, boundary_conditions, **kwargs)
    from transiflow.interface import create as ref_create
    return ref_create(parameters, nx, ny, nz, dim, dof, x, y, z, boundary_conditions, backend=backend)
