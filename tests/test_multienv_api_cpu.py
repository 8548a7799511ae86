"""The reference-shaped single-env surface (`OvercookedMultiEnv`) on the CPU emulation of the device code: the same
cases the GPU suite runs through the CUDA library (tests/multienv_cases.py) -- golden traces of the live reference,
the SURVEY A.7 known answer, the MultiAgentEnv step / reset protocol with an embedded partner."""
import pytest

from tests import multienv_cases as cases
from tests.parity_util import EmuMultiEnv


def _kw():
    return dict(device="cpu", env_cls=EmuMultiEnv)


@pytest.mark.parametrize("name", cases.TRACES)
def test_multi_step_matches_reference_trace(name):
    cases.run_multi_step_matches_reference_trace(name, **_kw())


def test_known_answer_survey_a7():
    cases.run_known_answer_survey_a7(**_kw())


def test_multiagentenv_step_reset_with_partner():
    cases.run_multiagentenv_step_reset_with_partner(**_kw())


def test_partner_selection_and_n_step():
    cases.run_partner_selection_and_n_step(**_kw())
