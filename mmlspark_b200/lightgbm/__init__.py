"""Host-side mirror of com.microsoft.ml.spark.lightgbm (estimators, TrainParams wire format, trainCore,
rendezvous) on top of the B200 engine's C ABI.  Same class / param / method names as the reference's PySpark
wrappers (lightgbm/src/main/python/mmlspark/lightgbm/*.py)."""
from .estimators import (Frame, LightGBMBooster, LightGBMClassificationModel, LightGBMClassifier, LightGBMRanker,
                         LightGBMRankerModel, LightGBMRegressionModel, LightGBMRegressor)
from .params import TrainParams, dataset_params, scala_double
from .train_utils import DriverRendezvous, LightGBMDelegate, count_cardinality, train_core

__all__ = ["Frame", "LightGBMBooster", "LightGBMClassificationModel", "LightGBMClassifier", "LightGBMRanker", "LightGBMRankerModel",
           "LightGBMRegressionModel", "LightGBMRegressor", "TrainParams", "dataset_params", "scala_double", "DriverRendezvous",
           "LightGBMDelegate", "count_cardinality", "train_core"]
