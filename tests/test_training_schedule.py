"""CPU: host logic of the offline training driver."""
import pytest


def test_epsilon_schedule_is_the_reference_formula():
    """training_pipeline.py:320: epsilon = max(0.01, 0.1 - episode_num / 5000)."""
    from marllb_b200.training_pipeline import MAX_EPISODE_STEPS, epsilon_schedule
    assert epsilon_schedule(0) == pytest.approx(0.1)
    assert epsilon_schedule(250) == pytest.approx(0.05)
    assert epsilon_schedule(449) == pytest.approx(0.1 - 449 / 5000.0)
    assert epsilon_schedule(450) == pytest.approx(0.01)          # 0.1 - 0.09: the floor is reached at episode 450
    assert epsilon_schedule(10 ** 6) == 0.01
    assert all(epsilon_schedule(e) >= epsilon_schedule(e + 1) for e in range(0, 600, 7))
    assert MAX_EPISODE_STEPS == 200                              # training_pipeline.py:322
