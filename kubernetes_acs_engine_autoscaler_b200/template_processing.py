"""ARM template surgery for a scale-out deployment (the step right after the hot path: it consumes the
`new_pool_sizes` that fulfill_pending produced; mirror of reference autoscaler/template_processing.py:5-263).

acs-engine templates describe a pool's NICs, VMs and VM extensions once, with a `copy` count.  An incremental
deployment that must create exactly the missing node indexes (a pool that lost node 1 of {0,1,2} needs "1" again,
not "3") cannot use the count, so the counted resource of every pool that grows is replaced by one concrete
resource per NEW index; pools that do not grow are dropped, as are the NSG and the outputs section.  CPU-side
JSON work; exact behaviour (insertion order, error messages, in-place mutation of the argument for the unroll_*
functions) follows the reference because deployments and its known-answer tests depend on it.
"""
import copy
import json

_NIC = "[concat(variables('{p}VMNamePrefix'), 'nic-', copyIndex(variables('{p}Offset')))]"
_STORAGE = ("[concat(variables('storageAccountPrefixes')[mod(add(copyIndex(),variables('{p}StorageAccountOffset')),"
            "variables('storageAccountPrefixesCount'))],variables('storageAccountPrefixes')[div(add(copyIndex(),"
            "variables('{p}StorageAccountOffset')),variables('storageAccountPrefixesCount'))],variables('{p}AccountName'))]")
_AVSET = "[variables('{p}AvailabilitySet')]"
_VM = "[concat(variables('{p}VMNamePrefix'), copyIndex(variables('{p}Offset')))]"
_EXT = "[concat(variables('{p}VMNamePrefix'), copyIndex(variables('{p}Offset')),'/cse', copyIndex(variables('{p}Offset')))]"
_NSG_DEP = "[concat('Microsoft.Network/networkSecurityGroups/', variables('nsgName'))]"


def get_new_nodes_indexes(pool, new_pool_size):
    """the lowest free node indexes that bring `pool` to `new_pool_size` (template_processing.py:235-252):
    a pool holding nodes {2, 4} scaled to 5 gets 0, 1 and 3."""
    wanted = new_pool_size - pool.actual_capacity
    indexes, idx = [], 0
    while len(indexes) < wanted:
        if not pool.has_node_with_index(idx):
            indexes.append(idx)
        idx += 1
    return indexes


def _unroll(template, pool, new_pool_size, is_counted, concrete_name, what, substitute_index):
    """replace the counted resource `is_counted` selects by one copy per new node index (each inserted at the
    FRONT of the resource list, so the highest index ends up first, as upstream)."""
    resources = template['resources']
    counted = None
    for i, res in enumerate(resources):
        if is_counted(res['name']):
            counted = copy.deepcopy(resources.pop(i))
            break
    if not counted:
        raise ValueError('Could not find the %s resource for the specified agent pool' % what)
    offset_expr = "copyIndex(variables('{}Offset'))".format(pool.name)
    for index in get_new_nodes_indexes(pool, new_pool_size):
        one = copy.deepcopy(counted)
        one.pop('copy')
        one['name'] = concrete_name(index)
        if substitute_index:  # every copyIndex(...) of the pool inside the resource becomes the literal index
            one = json.loads(json.dumps(one).replace(offset_expr, str(index)))
        resources.insert(0, one)
    return template


def unroll_nic(template, pool, new_pool_size):
    prefix = "[concat(variables('{}VMNamePrefix'), 'nic-'".format(pool.name)
    return _unroll(template, pool, new_pool_size, lambda name: name.startswith(prefix),
                   lambda i: "[concat(variables('{}VMNamePrefix'), 'nic-', {})]".format(pool.name, i), 'NIC', False)


def unroll_vm(template, pool, new_pool_size):
    counted = _VM.format(p=pool.name)
    return _unroll(template, pool, new_pool_size, lambda name: name == counted,
                   lambda i: "[concat(variables('{}VMNamePrefix'), {})]".format(pool.name, i), 'virtualMachines', True)


def unroll_vm_extension(template, pool, new_pool_size):
    counted = _EXT.format(p=pool.name)
    return _unroll(template, pool, new_pool_size, lambda name: name == counted,
                   lambda i: "[concat(variables('{}VMNamePrefix'), {},'/cse', {})]".format(pool.name, i, i),
                   'virtualMachines/extensions', True)


def unroll_resources(template, pools, new_pool_sizes):
    """NICs, VMs and VM extensions of every growing pool; storage accounts keep their count."""
    for pool in pools:
        size = new_pool_sizes[pool.name]
        if pool.actual_capacity == size:
            continue
        template = unroll_nic(template, pool, size)
        template = unroll_vm(template, pool, size)
        template = unroll_vm_extension(template, pool, size)
    return template


def delete_resources_by_name(template, names):
    template['resources'][:] = [r for r in template['resources'] if r['name'] not in names]
    return template


def delete_unchanged_pools(template, unchanged_pools):
    names = {tpl.format(p=pool.name) for pool in unchanged_pools for tpl in (_NIC, _STORAGE, _AVSET, _VM, _EXT)}
    return delete_resources_by_name(template, names)


def delete_nsg(template):
    """drop the network security group and every dependency on it (returns a copy)."""
    template = copy.deepcopy(template)
    resources = template['resources']
    nsg_index = -1  # (sic) no NSG -> the LAST resource goes, exactly as upstream's pop(-1)
    for i, res in enumerate(resources):
        kind = res['type']
        if kind == 'Microsoft.Network/networkSecurityGroups':
            nsg_index = i
        dep = {'Microsoft.Network/virtualNetworks': _NSG_DEP, 'Microsoft.Network/networkInterfaces': "[variables('nsgID')]",
               'Microsoft.Network/loadBalancers': "[variables('nsgID')]"}.get(kind)
        if dep is not None and dep in res['dependsOn']:
            res['dependsOn'].remove(dep)  # first occurrence only
    resources.pop(nsg_index)
    return template


def delete_outputs_section(template):
    template.pop('outputs')
    return template


def delete_master_vm_extension(template):
    """used once at login (cluster.py:109).  (sic) when the master extension is absent the last resource is
    removed: upstream pops the loop variable, not the index it found."""
    target = _EXT.format(p='master')
    resources = template['resources']
    index = len(resources) - 1
    for i, res in enumerate(resources):
        if res['name'] == target:
            index = i
            break
    resources.pop(index)
    return template


def prepare_template_for_scale_out(template, pools, new_pool_sizes):
    """the template an incremental deployment of `new_pool_sizes` submits (template_processing.py:115-135)."""
    growing = [p for p in pools if p.actual_capacity < new_pool_sizes[p.name]]
    unchanged = [p for p in pools if not p.actual_capacity < new_pool_sizes[p.name]]
    template = delete_nsg(copy.deepcopy(template))
    template = delete_unchanged_pools(template, unchanged)
    template = unroll_resources(template, growing, new_pool_sizes)
    return delete_outputs_section(template)
