'''Host-side surface of the backend that does not need a device: the parameter file format of the reference
(/root/reference/transiflow/interface/BaseInterface.py:10-29,118-215) and the method set callers rely on.'''
import json

import numpy

from transiflow_b200 import Interface
from transiflow_b200.interface import DeviceMatrix, ParameterEncoder, parameter_decoder

# the methods Continuation / TimeIntegration / JaDa / user scripts call on an interface
# (BaseInterface.py:79-292, docs/custom-backend.rst:10-32)
BASE_INTERFACE_METHODS = [
    'vector', 'vector_from_array', 'array_from_vector', 'set_parameter', 'get_parameter', 'save_json', 'save_parameters',
    'save_state', 'load_json', 'load_parameters', 'load_state', 'rhs', 'jacobian', 'mass_matrix', 'solve', 'eigs',
    '_debug_print', '_debug_print_residual',
]


def test_interface_has_the_base_interface_methods():
    for name in BASE_INTERFACE_METHODS:
        assert callable(getattr(Interface, name, None)), name


def test_parameter_files_round_trip_numpy_and_complex_values():
    params = {'Eigenvalue Solver': {'Target': 1 + 3j, 'Number of Eigenvalues': numpy.int64(5)},
              'Reynolds Number': numpy.float64(100.0), 'Profile': numpy.arange(3.0), 'c64': numpy.complex64(2 - 1j)}
    text = json.dumps(params, cls=ParameterEncoder)
    # the reference's on-disk convention for complex numbers
    assert json.loads(text)['Eigenvalue Solver']['Target'] == {'__complex__': True, 'real': 1.0, 'imag': 3.0}
    back = json.loads(text, object_hook=parameter_decoder)
    assert back['Eigenvalue Solver']['Target'] == 1 + 3j
    assert back['Eigenvalue Solver']['Number of Eigenvalues'] == 5 and isinstance(back['Eigenvalue Solver']['Number of Eigenvalues'], int)
    assert back['Reynolds Number'] == 100.0 and back['Profile'] == [0.0, 1.0, 2.0] and back['c64'] == 2 - 1j


def test_device_matrix_exposes_what_the_eigen_solver_glue_reads():
    # JaDa.Op reads mat.data.dtype, mat.shape, mat.dtype and uses mat @ x (JaDa.py:24-34)
    for name in ('data', 'tocsr', 'tocsc', '__matmul__', '__sub__', '__rmul__'):
        assert hasattr(DeviceMatrix, name), name
