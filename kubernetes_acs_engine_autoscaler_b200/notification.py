"""Scale / drain notifications (same interface and message text as reference
autoscaler/notification.py:55-156).

What the path hands over is large at synthetic scale -- `notify_scale(new_pool_sizes, pods, current_pool_sizes)`
gets every pending pod -- so this notifier aggregates: the chat message already names at most four pods ("a, b, c,
d, and N others", notification.py:33-40), and the per-pod structured log lines (notification.py:43-52) are
emitted only up to `detail_limit` pods, followed by one summary record.  The HTTP POST is the adapter boundary:
`post(url, json=payload)` defaults to `requests.post` when that package is importable.
"""
import json
import logging

logger = logging.getLogger('autoscaler.notification')
struct_logger = logging.getLogger('autoscaler.notification.struct')

USERNAME = "kubernetes-acs-engine-autoscaler"


def pod_string(pods):
    """'ns/a, ns/b' -- or the first four and a count when there are more than five (notification.py:33-40)."""
    pods = list(pods)
    names = ['{}/{}'.format(p.namespace, p.name) for p in pods[:5 if len(pods) <= 5 else 4]]
    if len(pods) > 5:
        return '{}, and {} others'.format(', '.join(names), len(pods) - 4)
    return ', '.join(names)


def struct_log(message, pods, extra=None, detail_limit=None):
    """one structured record per pod, as upstream; beyond `detail_limit` pods a single summary record instead of
    the remaining ones (the reference would emit millions of lines on a synthetic-scale tick)."""
    pods = list(pods)
    shown = pods if detail_limit is None else pods[:detail_limit]
    for pod in shown:
        record = {'pod_name': '{}/{}'.format(pod.namespace, pod.name), 'pod_id': pod.uid,
                  '_log_streaming_target_mapping': USERNAME}
        if extra:
            record.update(extra)
        struct_logger.debug(message, extra=record)
    if len(shown) < len(pods):
        record = {'pods_not_listed': len(pods) - len(shown), '_log_streaming_target_mapping': USERNAME}
        if extra:
            record.update(extra)
        struct_logger.debug(message + ' (summary)', extra=record)


def _default_post(url, json=None):
    import requests
    return requests.post(url, json=json)


class Notifier(object):
    def __init__(self, hook=None, post=None, detail_limit=1000):
        self.hook = hook
        self.post = post or _default_post
        self.detail_limit = detail_limit

    def _send(self, message, username=USERNAME):
        if not self.hook:
            logger.debug('SLACK_HOOK not configured.')
            return
        try:
            resp = self.post(self.hook, json={"text": message, "username": username, "icon_emoji": ":camel:"})
            logger.debug('SLACK: %s', getattr(resp, 'text', resp))
        except Exception as e:  # requests.exceptions.ConnectionError upstream; never let a chat hook fail a tick
            logger.critical('Failed to SLACK: %s', e)

    def notify_scale(self, units_requested, pods, units_actual):
        struct_log('scale', pods, extra={'units_requested': units_requested}, detail_limit=self.detail_limit)
        self._send('Scaled up from {} to new capacity {}\nChange triggered by {}'.format(
            units_actual, units_requested, pod_string(pods)) if self.hook else None)

    def notify_failed_to_scale(self, selectors_hash, pods):
        struct_log('failed to scale', pods, extra={'selectors_hash': selectors_hash}, detail_limit=self.detail_limit)
        self._send('Failed to scale {} sufficiently. Backing off...\nPods affected: {}'.format(
            json.dumps(selectors_hash), pod_string(pods)) if self.hook else None, username="kubernetes-acs-enginbe-autoscaler")

    def notify_invalid_pod_capacity(self, pod, recommended_capacity):
        struct_log('invalid pod capacity', [pod], extra={'recommended_capacity': str(recommended_capacity)})
        self._send(("Pending pod {}/{} cannot fit {}. Please check that requested resource amount is consistent with "
                    "node selectors (recommended max: {}). Scheduling skipped.").format(
            pod.namespace, pod.name, json.dumps(pod.selectors), recommended_capacity) if self.hook else None)

    def notify_drained_node(self, node, pods):
        struct_log('drain', pods, extra={'node': str(node)}, detail_limit=self.detail_limit)
        self._send('Node {} drained.\nPod affected: {}'.format(node, pod_string(pods)) if self.hook else None)
