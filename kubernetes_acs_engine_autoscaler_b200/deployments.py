"""Single-flight guard around ARM deployments (same contract as reference autoscaler/deployments.py):
skip when one is still running or when the requested sizes did not change."""
import logging

logger = logging.getLogger('autoscaler.deployments')


class Deployments(object):
    def __init__(self):
        self._current_deployment = None
        self.requested_pool_sizes = None

    def deploy(self, func, new_pool_sizes):
        running = self._current_deployment is not None and not self._current_deployment.done()
        if running:
            logger.info('Another deployment is already in progress')
            return
        if self.requested_pool_sizes and self.requested_pool_sizes == new_pool_sizes:
            logger.info('Requested a new deployment with unchanged pool sizes, skipping.')
            return
        self.requested_pool_sizes = new_pool_sizes
        self._current_deployment = func()
        if hasattr(self._current_deployment, 'wait') and hasattr(self._current_deployment, 'result'):
            self._current_deployment.wait()
            logger.info('Deployment finished: {}'.format(self._current_deployment.result()))
