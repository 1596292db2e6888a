"""CPU-side checks of the pybind11 `GPP` module (the reference's `moe.build.GPP` surface, gpp_python*.cpp): every name the
reference module exports exists, the parameter / randomness containers behave like the reference's, off-path entries
and device-less calls fail with the library's own exception classes.  No GPU needed."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cornell-moe_b200"))

import GPP as C_GP  # noqa: E402

# boost::python::def / class_ names in gpp_python*.cpp of the reference
REFERENCE_FUNCTIONS = """compute_expected_improvement compute_grad_expected_improvement
multistart_expected_improvement_optimization evaluate_EI_at_point_list heuristic_expected_improvement_optimization
compute_expected_improvement_mcmc compute_grad_expected_improvement_mcmc multistart_expected_improvement_mcmc_optimization
evaluate_EI_mcmc_at_point_list compute_posterior_mean compute_grad_posterior_mean compute_knowledge_gradient
compute_grad_knowledge_gradient multistart_knowledge_gradient_optimization posterior_mean_optimization
evaluate_KG_at_point_list compute_knowledge_gradient_mcmc compute_grad_knowledge_gradient_mcmc
multistart_knowledge_gradient_mcmc_optimization evaluate_KG_mcmc_at_point_list compute_log_likelihood
compute_hyperparameter_grad_log_likelihood multistart_hyperparameter_optimization restarted_hyperparameter_optimization
evaluate_log_likelihood_at_hyperparameter_list run_cpp_tests""".split()
REFERENCE_CLASSES = ["GaussianProcess", "GaussianProcessMCMC", "GradientDescentParameters", "NewtonParameters",
                     "RandomnessSourceContainer", "OptimizerTypes", "DomainTypes", "LogLikelihoodTypes",
                     "OptimalLearningException", "BoundsException", "InvalidValueException", "SingularMatrixException"]
GP_METHODS = """compute_mean_of_points compute_mean_of_additional_points compute_grad_mean_of_points
compute_variance_of_points compute_cholesky_variance_of_points compute_grad_variance_of_points
compute_grad_cholesky_variance_of_points add_sampled_points sample_point_from_gp sample_global_optima set_explicit_seed
set_randomized_seed reset_to_most_recent_seed print_historical_data""".split()
RNG_METHODS = """SetExplicitUniformGeneratorSeed SetRandomizedUniformGeneratorSeed ResetUniformRNGSeed
SetExplicitNormalRNGSeed SetRandomizedNormalRNGSeed SetNormalRNGSeedPythonList ResetNormalRNGSeed PrintState""".split()


def test_every_reference_name_exists():
    missing = [n for n in REFERENCE_FUNCTIONS + REFERENCE_CLASSES if not hasattr(C_GP, n)]
    assert not missing, missing
    assert not [m for m in GP_METHODS if not hasattr(C_GP.GaussianProcess, m)]
    assert not [m for m in RNG_METHODS if not hasattr(C_GP.RandomnessSourceContainer, m)]


def test_exception_hierarchy_and_enums():
    # gpp_python.cpp:189-206: everything derives from OptimalLearningException
    for name in ("BoundsException", "InvalidValueException", "SingularMatrixException"):
        assert issubclass(getattr(C_GP, name), C_GP.OptimalLearningException)
    assert issubclass(C_GP.OptimalLearningException, Exception)
    assert {C_GP.OptimizerTypes.null, C_GP.OptimizerTypes.gradient_descent, C_GP.OptimizerTypes.newton}
    assert {C_GP.DomainTypes.tensor_product, C_GP.DomainTypes.simplex}
    assert {C_GP.LogLikelihoodTypes.log_marginal_likelihood, C_GP.LogLikelihoodTypes.leave_one_out_log_likelihood}


def test_parameter_structs_take_the_reference_positional_arguments():
    # gpp_optimizer_parameters.hpp:46-71 / :120-160
    C_GP.GradientDescentParameters(200, 50, 2, 4, 0.7, 1.0, 0.5, 1.0e-10)
    C_GP.NewtonParameters(10, 100, 1.01, 1.0e-2, 1.0, 1.0e-9)
    with pytest.raises(TypeError):
        C_GP.GradientDescentParameters(1, 2, 3)


def test_randomness_container():
    r = C_GP.RandomnessSourceContainer(4)
    r.SetExplicitUniformGeneratorSeed(314)
    r.SetExplicitNormalRNGSeed(100)
    assert r.SetNormalRNGSeedPythonList([7, 8, 9, 10], [1, 0, 1, 0]) is True
    assert r.SetNormalRNGSeedPythonList([1, 2], [1, 1]) is False  # wrong length, as in gpp_python_common.cpp
    r.SetRandomizedNormalRNGSeed(5)
    r.SetRandomizedUniformGeneratorSeed(5)
    r.ResetUniformRNGSeed()
    r.ResetNormalRNGSeed()


def test_off_path_entries_raise_the_library_exception():
    for name in ("restarted_hyperparameter_optimization", "multistart_hyperparameter_optimization",
                 "heuristic_expected_improvement_optimization", "run_cpp_tests"):
        with pytest.raises(C_GP.OptimalLearningException):
            getattr(C_GP, name)() if name == "run_cpp_tests" else getattr(C_GP, name)(1, 2, x=3)


def test_no_device_no_compute():
    if C_GP.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(C_GP.OptimalLearningException):
        C_GP.GaussianProcess([1.0, [1.0, 1.0]], [0.1, 0.2, 0.7, 0.4], [0.3, 0.9], [0.01], [], 0, 2, 2)
    with pytest.raises(C_GP.OptimalLearningException):
        C_GP.GaussianProcessMCMC([1.0, 1.0, 1.0], [0.01], [0.1, 0.2, 0.7, 0.4], [0.3, 0.9], [], 1, 0, 2, 2)
    with pytest.raises(C_GP.OptimalLearningException):  # positional signature of gpp_python_model_selection.cpp:43-50
        C_GP.compute_log_likelihood([0.1, 0.2, 0.7, 0.4], [0.3, 0.9], 2, 2, C_GP.LogLikelihoodTypes.log_marginal_likelihood,
                                    [1.0, [1.0, 1.0]], [], 0, [0.01])
    # short input lists are rejected before any device work (BoundsException, like CopyPylistToVector's size check)
    with pytest.raises(C_GP.BoundsException):
        C_GP.GaussianProcess([1.0, [1.0, 1.0]], [0.1, 0.2], [0.3, 0.9], [0.01], [], 0, 2, 2)
