"""ModelFreeBasePolicy (blackbox_mpc/policies/model_free_base_policy.py:1-36): `act(observations, t)` returns
actions only (no predicted observation / reward), which is how utils/rollouts.py tells the two kinds apart."""


class ModelFreeBasePolicy:
    def act(self, observations, t):
        raise Exception("act function is not implemented yet")

    def reset(self):
        raise Exception("reset function is not implemented yet")
