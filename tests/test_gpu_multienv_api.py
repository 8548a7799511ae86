"""GPU: the reference-shaped single-env surface (`OvercookedMultiEnv.multi_step / multi_reset /
get_observation2`, gym_comm/envs/overcooked_env.py:105-297) against golden traces of the live
reference (cases in tests/multienv_cases.py), through the CUDA library."""
import pytest

from tests import multienv_cases as cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", cases.TRACES)
def test_multi_step_matches_reference_trace(name):
    cases.run_multi_step_matches_reference_trace(name)


def test_known_answer_survey_a7():
    cases.run_known_answer_survey_a7()


def test_multiagentenv_step_reset_with_partner():
    cases.run_multiagentenv_step_reset_with_partner()


def test_partner_selection_and_n_step():
    cases.run_partner_selection_and_n_step()
