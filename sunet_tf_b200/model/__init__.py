"""Import-path mirror of the reference's ``model`` package: ``from sunet_tf_b200.model.SUNet import SUNet_model``."""
