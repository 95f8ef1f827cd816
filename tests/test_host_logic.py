"""Host-side mirror of the reference's kernel DSL / Params (no GPU needed): hyperparameter layout,
bounds, flattening into the C-ABI term list -- checked against the oracle's restatement."""
import numpy as np
import pytest

import oracle
import spark_gp_b200 as sg
from spark_gp_b200 import _native as N


def _pairs():
    return [
        (lambda: 1 * sg.ARDRBFKernel(5) + sg.const(1) * sg.EyeKernel(),
         lambda: 1 * oracle.ARDRBFKernel(5) + oracle.const(1) * oracle.EyeKernel()),
        (lambda: sg.Scalar(1.0).between(0).and_(30) * sg.RBFKernel(0.1, 1e-6, 10) + sg.WhiteNoiseKernel(0.5, 0, 1),
         lambda: oracle.Scalar(1.0).between(0).and_(30) * oracle.RBFKernel(0.1, 1e-6, 10) + oracle.WhiteNoiseKernel(0.5, 0, 1)),
        (lambda: sg.RBFKernel(10), lambda: oracle.RBFKernel(10)),
        (lambda: 2.0 * (1.5 * sg.ARDRBFKernel(np.array([0.2, 0.3])) + sg.const(0.5) * sg.EyeKernel()),
         lambda: 2.0 * (1.5 * oracle.ARDRBFKernel(np.array([0.2, 0.3])) + oracle.const(0.5) * oracle.EyeKernel())),
    ]


@pytest.mark.parametrize("idx", range(4))
def test_dsl_matches_oracle(idx):
    mk, mo = _pairs()[idx]
    k, o = mk(), mo()
    assert np.allclose(k.getHyperparameters(), o.get_hyperparameters())
    assert k.numberOfHyperparameters() == o.number_of_hyperparameters()
    for a, b in zip(k.hyperparameterBoundaries(), o.hyperparameter_boundaries()):
        assert np.array_equal(a, b)
    assert np.isclose(k.whiteNoiseVar, o.white_noise_var)
    theta = k.getHyperparameters() * 1.7 + 0.01
    k.setHyperparameters(theta); o.set_hyperparameters(theta)
    assert np.allclose(k.getHyperparameters(), o.get_hyperparameters())
    assert np.isclose(k.whiteNoiseVar, o.white_noise_var)
    assert str(k) == str(o)
    # self kernel = sum of every leaf's scale (each leaf has k(x,x)=1)
    assert np.isclose(sum(t["scale"] for t in k.flatten()), o.self_kernel(np.zeros(5)))


def test_flatten_scales_multiply_through_nesting():
    k = 2.0 * (1.5 * sg.ARDRBFKernel(np.array([0.2, 0.3])) + sg.const(0.5) * sg.EyeKernel())
    t = k.flatten()
    assert [x["type"] for x in t] == [N.SGP_TERM_ARD, N.SGP_TERM_EYE]
    assert np.isclose(t[0]["scale"], 3.0) and np.isclose(t[1]["scale"], 1.0)
    assert np.allclose(k.getHyperparameters(), [2.0, 1.5, 0.2, 0.3])


def test_params_defaults_and_getkernel():
    gp = sg.GaussianProcessRegression()
    assert (gp._datasetSizeForExpert, gp._activeSetSize, gp._sigma2, gp._maxIter, gp._tol) == (100, 100, 1e-3, 100, 1e-6)
    k = gp.setSigma2(1e-4).setKernel(lambda: 1 * sg.ARDRBFKernel(5) + sg.const(1) * sg.EyeKernel()).getKernel()
    assert np.isclose(k.whiteNoiseVar, 1.0001)            # GPC:18 appends sigma2.const * Eye
    assert len(k.getHyperparameters()) == 6
    with pytest.raises(ValueError):
        sg.Scalar(-1.0) * sg.RBFKernel()
    with pytest.raises(ValueError):
        sg.Scalar(1.0, 2.0, 1.0)
