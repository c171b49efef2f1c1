"""The plugin boundary: same abstract interface as the reference's ``gdmix.models.api.Model``
(gdmix-trainer/src/gdmix/models/api.py:4-84)."""
import abc


class Model(abc.ABC):
    """Every model the drivers run exposes train / predict / export and parses its own parameters."""

    def __init__(self, raw_model_params):
        self.raw_model_params = raw_model_params

    @abc.abstractmethod
    def train(self, training_data_dir, validation_data_dir, metadata_file, checkpoint_path, execution_context,
              schema_params):
        raise NotImplementedError("Must be implemented in subclasses.")

    @abc.abstractmethod
    def predict(self, output_dir, input_data_path, metadata_file, checkpoint_path, execution_context, schema_params):
        raise NotImplementedError("Must be implemented in subclasses.")

    @abc.abstractmethod
    def export(self, output_model_dir):
        raise NotImplementedError("Must be implemented in subclasses.")

    @abc.abstractmethod
    def _parse_parameters(self, raw_model_parameters):
        raise NotImplementedError("Must be implemented in subclasses.")
