"""Agent pool = name + instance type + its nodes (mirror of reference autoscaler/agent_pool.py)."""
import logging

from .capacity import get_capacity_for_instance_type

logger = logging.getLogger('autoscaler.agent_pool')


class AgentPool(object):
    def __init__(self, pool_name, instance_type, nodes):
        self.name = pool_name
        self.nodes = nodes
        self.unschedulable_nodes = [n for n in nodes if n.unschedulable]
        self.max_size = 100  # ACS limit (agent_pool.py:18); a plain attribute, callers may raise it
        self.instance_type = instance_type

    @property
    def actual_capacity(self):
        return len(self.nodes)

    @property
    def unit_capacity(self):
        return get_capacity_for_instance_type(self.instance_type)

    def reclaim_unschedulable_nodes(self, new_desired_capacity):
        """uncordon just enough of our own cordoned nodes before asking for new VMs
        (agent_pool.py:30-44; the uncordon itself is the kube adapter)."""
        desired = min(self.max_size, new_desired_capacity)
        schedulable = self.actual_capacity - len(self.unschedulable_nodes)
        if schedulable >= desired:
            return
        for node in self.unschedulable_nodes:
            if node.uncordon():
                schedulable += 1
                if schedulable == desired:
                    break

    def has_node_with_index(self, index):
        return any(node.index == index for node in self.nodes)
