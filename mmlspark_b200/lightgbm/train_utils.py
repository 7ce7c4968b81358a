"""Per-task training runtime: driver rendezvous, network init, createBooster, trainCore (early stopping, delegate
hooks).  Mirrors lightgbm/src/main/scala/com/microsoft/ml/spark/lightgbm/TrainUtils.scala and the driver side of
LightGBMBase.scala:392-430 / LightGBMUtils.scala:59-104.  One rank-thread per partition == one Spark task thread
of the reference's local mode; the thread's CUDA device replaces the executor's cores."""
import logging
import socket
import threading
import time

import numpy as np

from .. import capi
from .params import scala_double

log = logging.getLogger("mmlspark_b200.lightgbm")

# LightGBMConstants.scala
DEFAULT_LOCAL_LISTEN_PORT = 12400
DEFAULT_LISTEN_TIMEOUT = 120
MAX_PORT = 65535
IGNORE_STATUS = "ignore"
FINISHED_STATUS = "finished"
NETWORK_RETRIES = 3
INITIAL_DELAY_MS = 1000


class DriverRendezvous:
    """createDriverNodesThread (LightGBMBase.scala:392-430): accept numTasks connections, each sends `host:port`
    (or `ignore` for an empty partition); then the comma-joined list goes back to every live task."""

    def __init__(self, num_tasks, listen_port=0, timeout=1200.0, host="127.0.0.1"):
        self.num_tasks = num_tasks
        self.server = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.server.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.server.bind((host, listen_port))
        self.server.listen(max(num_tasks, 1))
        self.server.settimeout(timeout)
        self.host = host
        self.port = self.server.getsockname()[1]
        self.error = None
        self.nodes = None
        self.thread = threading.Thread(target=self._run, daemon=True)

    def start(self):
        self.thread.start()
        return self.host, self.port

    def _run(self):
        try:
            conns, empty = [], 0
            while len(conns) + empty < self.num_tasks:
                c, _ = self.server.accept()
                f = c.makefile("r")
                comm = f.readline().rstrip("\n")
                if comm == IGNORE_STATUS:
                    empty += 1
                    c.close()
                elif comm == FINISHED_STATUS:
                    c.close()
                    break
                else:
                    conns.append((c, comm))
            self.nodes = ",".join(comm for _, comm in conns)
            for c, _ in conns:
                c.sendall((self.nodes + "\n").encode())
            for c, _ in conns:
                c.close()
        except Exception as e:   # noqa
            self.error = e
        finally:
            self.server.close()

    def join(self, timeout=None):
        self.thread.join(timeout)
        if self.error is not None:
            raise self.error


def find_open_port(default_listen_port, worker_id, num_tasks_per_exec=1):
    """TrainUtils.findOpenPort (:193-220): first free port at defaultListenPort + workerId * numTasksPerExec."""
    base = default_listen_port + worker_id * num_tasks_per_exec
    if base > MAX_PORT:
        raise RuntimeError("Error: port %d out of range, possibly due to too many executors or unknown error" % base)
    port = base
    while True:
        s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        try:
            s.bind(("", port))
            return s, port
        except OSError:
            s.close()
            port += 1
            if port > MAX_PORT:
                raise RuntimeError("Error: port %d out of range, possibly due to networking or firewall issues" % base)
            if port - base > 1000:
                raise RuntimeError("Error: Could not find open port after 1k tries")


def get_network_init_nodes(driver_host, driver_port, local_listen_port, ignore_task, local_host="127.0.0.1"):
    """TrainUtils.getNetworkInitNodes (:236-277)"""
    with socket.create_connection((driver_host, driver_port)) as s:
        status = IGNORE_STATUS if ignore_task else "%s:%d" % (local_host, local_listen_port)
        s.sendall((status + "\n").encode())
        if ignore_task:
            return IGNORE_STATUS
        return s.makefile("r").readline().rstrip("\n")


def network_init(nodes, local_listen_port, retry=NETWORK_RETRIES, delay_ms=INITIAL_DELAY_MS):
    """TrainUtils.networkInit (:279-295): LGBM_NetworkInit with exponential back-off retries."""
    try:
        capi.network_init(nodes, local_listen_port, DEFAULT_LISTEN_TIMEOUT, len(nodes.split(",")))
    except Exception as ex:   # noqa
        log.info("NetworkInit failed with exception on local port %d with exception: %s", local_listen_port, ex)
        time.sleep(delay_ms / 1000.0)
        if retry > 0:
            network_init(nodes, local_listen_port, retry - 1, delay_ms * 2)
        else:
            raise


def get_main_worker_port(nodes):
    """TrainUtils.getMainWorkerPort (:302-316): the first node returns the model."""
    first = nodes.split(",")[0]
    hp = first.split(":")
    if len(hp) != 2:
        raise RuntimeError("Error: could not parse main worker host and port correctly")
    return int(hp[1])


def count_cardinality(ids):
    """DatasetUtils.countCardinality (dataset/DatasetUtils.scala:18-40): run lengths of consecutive equal group ids."""
    out = []
    prev, run = object(), 0
    for v in ids:
        if run and v == prev:
            run += 1
        else:
            if run:
                out.append(run)
            prev, run = v, 1
    if run:
        out.append(run)
    return out


class LightGBMDelegate:
    """LightGBMDelegate.scala: user hooks.  Override what you need."""

    def beforeTrainBatch(self, batchIndex, dataset, model): pass
    def afterTrainBatch(self, batchIndex, dataset, model): pass
    def beforeGenerateTrainDataset(self, batchIndex, partitionId, trainParams): pass
    def afterGenerateTrainDataset(self, batchIndex, partitionId, trainParams): pass
    def beforeTrainIteration(self, batchIndex, partitionId, curIters, trainParams, booster, hasValid): pass
    def afterTrainIteration(self, batchIndex, partitionId, curIters, trainParams, booster, hasValid, isFinished, trainEvalResults, validEvalResults): pass
    def getLearningRate(self, batchIndex, partitionId, curIters, trainParams, previousLearningRate): return previousLearningRate


def create_booster(train_params, train_ds, valid_ds):
    """TrainUtils.createBooster (:25-37)"""
    booster = capi.Booster(train_ds, train_params.to_string())
    if train_params.modelString:
        other = capi.Booster(model_str=train_params.modelString)
        try:
            booster.merge(other)
        finally:
            other.free()
    if valid_ds is not None:
        booster.add_valid(valid_ds)
    return booster


def update_one_iteration(train_params, booster, iters, train_label=None, classification=False):
    """TrainUtils.updateOneIteration (:67-90): any native exception is logged and turned into isFinished=true."""
    try:
        fobj = train_params.objectiveParams.fobj
        if fobj is not None:
            preds = booster.get_predict(0)                       # innerPredict(0, classification)
            grad, hess = fobj.getGradient(preds, train_label)
            fin = booster.update_one_iter_custom(np.asarray(grad, dtype=np.float32), np.asarray(hess, dtype=np.float32))
        else:
            fin = booster.update_one_iter()
        log.info("LightGBM running iteration: %d with is finished: %s", iters, fin)
        return fin
    except Exception as e:   # noqa
        log.warning("LightGBM reached early termination on one task, stopping training on task. This message should rarely occur. Inner exception: %s", e)
        return True


def train_core(batch_index, partition_id, train_params, booster, has_valid, train_label=None):
    """TrainUtils.trainCore (:92-159), including its quirks: the metric direction is decided by
    name prefix; with earlyStoppingRound == 0 the first non-improving iteration stops training; the returned best
    iteration is the 0-based index."""
    is_finished = False
    iters = 0
    eval_names = booster.eval_names()
    n = len(eval_names)
    best_score = [0.0] * n
    best_scores = [None] * n
    best_iter = [0] * n
    learning_rate = train_params.learningRate
    best_iter_result = None
    delegate = train_params.delegate
    while not is_finished and iters < train_params.numIterations:
        if delegate is not None:
            delegate.beforeTrainIteration(batch_index, partition_id, iters, train_params, booster, has_valid)
            new_lr = delegate.getLearningRate(batch_index, partition_id, iters, train_params, learning_rate)
            if new_lr != learning_rate:
                booster.reset_parameter("learning_rate=%s" % scala_double(new_lr))
                learning_rate = new_lr
        is_finished = update_one_iteration(train_params, booster, iters, train_label, train_params.kind == "classifier")
        train_eval = None
        if train_params.isProvideTrainingMetric and not is_finished:
            train_eval = dict(zip(eval_names, booster.get_eval(0).tolist()))
            for k, v in train_eval.items():
                log.info("Train %s=%s", k, v)
        valid_eval = None
        if has_valid and not is_finished:
            res = booster.get_eval(1).tolist()
            for index, (name, score) in enumerate(zip(eval_names, res)):
                log.info("Valid %s=%s", name, score)
                tol = train_params.improvementTolerance
                larger_better = name.startswith("auc") or name.startswith("ndcg@") or name.startswith("map@") or name.startswith("average_precision")
                better = (score - best_score[index] > tol) if larger_better else (score - best_score[index] < tol)
                if best_scores[index] is None or better:
                    best_score[index] = score
                    best_iter[index] = iters
                    best_scores[index] = list(res)
                elif iters - best_iter[index] >= train_params.earlyStoppingRound:
                    is_finished = True
                    log.info("Early stopping, best iteration is %d", best_iter[index])
                    best_iter_result = best_iter[index]
            valid_eval = dict(zip(eval_names, res))
        if delegate is not None:
            delegate.afterTrainIteration(batch_index, partition_id, iters, train_params, booster, has_valid, is_finished, train_eval, valid_eval)
        iters += 1
    return best_iter_result
