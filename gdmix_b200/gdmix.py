"""``python -m gdmix_b200.gdmix --k=v ...`` -- the command line the workflow DAG emits for every trainer job
(gdmix-workflow/src/gdmixworkflow/single_node/local_ops.py:15-23), same flags as ``python -m gdmix.gdmix``
(gdmix-trainer/src/gdmix/gdmix.py:13-36): unknown flags are ignored, a non-zero exit code means failure."""
import logging
import sys

from . import constants
from .drivers import DriverFactory
from .params import Params, SchemaParams

logging.basicConfig(level=logging.INFO)
logger = logging.getLogger(__name__)


def run(args):
    params = Params.__from_argv__(args, error_on_unknown=False)
    schema_params = SchemaParams.__from_argv__(args, error_on_unknown=False)
    logger.info(f"Parsed schema params amd gdmix args (params): {params}")
    driver = DriverFactory.get_driver(base_training_params=params, raw_model_params=args)
    if params.action == constants.ACTION_TRAIN:
        driver.run_training(schema_params=schema_params, export_model=True)
    elif params.action == constants.ACTION_INFERENCE:
        driver.run_inference(schema_params=schema_params)
    else:
        raise Exception(f"Unsupported action {params.action}")


if __name__ == "__main__":
    run(sys.argv)
