"""TEST / BASELINE INFRASTRUCTURE - not product code.

The UNMODIFIED reference (under /root/reference in the build container, or its staged copy baseline/_ref on the GPU
box), driven one frame at a time with the call sequence of its own driver, test/metrics_from_model.py:178-300 - JSON
strings in, python lists out - on CPU, with the dgl / pytransform3d import shims. This is bench.py's reference arm
(cpu_baseline.kind = "reference"); every number computed here comes out of the reference's functions.
"""
import json
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)


def available():
    from oracle import ref_env
    return ref_env.reference_available()


class ReferenceRunner:
    def __init__(self, config, gat_state, mlp_state, threads=1):
        import torch
        from oracle import ref_env
        torch.set_grad_enabled(False)
        torch.set_num_threads(threads)
        cwd = os.getcwd()
        self.parameters = ref_env.activate_config(config)
        import gat2, graph_generator, mlp as mlp_mod, pose_estimator_dataset_from_json as dataset        # noqa: E401
        import skeleton_matching_utils
        os.chdir(cwd)
        p = self.parameters
        self.gg, self.dataset, self.smu = graph_generator, dataset, skeleton_matching_utils
        n_feats = len(graph_generator.HumanGraphFromView.get_all_features('3'))
        # constructor arguments of the shipped models: train_skeleton_matching.py:40-56,148-149; metrics_from_model.py:90-100
        self.model = gat2.GAT2(None, 5, n_feats, 1, [40, 40, 40, 30], [10, 10, 8, 5], torch.nn.LeakyReLU(), torch.nn.Sigmoid(),
                               0., 0., 0.15, False, bias=True)
        self.model.load_state_dict({k: torch.as_tensor(v) for k, v in gat_state.items()})
        self.mlp = mlp_mod.PoseEstimatorMLP(input_dimensions=len(p.used_cameras) * len(p.joint_list) * p.numbers_per_joint,
                                            output_dimensions=54)
        self.mlp.load_state_dict({k: torch.as_tensor(v) for k, v in mlp_state.items()})
        self.torch = torch

    def run_frame(self, input_element):
        """metrics_from_model.py:178-300 for one frame; returns (proposals as {camera: head or None} dicts, list of [18 x 3] results)."""
        torch, p = self.torch, self.parameters
        processed_input = {}
        for cam in input_element:                                                   # :182-191
            data = json.loads(input_element[cam][0])
            cam_data = [s for s in data]
            if cam_data:
                processed_input[cam] = [json.dumps(cam_data), input_element[cam][1]]
        scenario = self.gg.MergedMultipleHumansDataset(processed_input, mode='test', limit=10000, debug=True,
                                                       alt=p.graph_alternative, verbose=False)
        if len(scenario.graphs) == 0:
            return [], []
        subgraph = scenario.graphs[0]
        indices = scenario.data['edge_nodes_indices'][0]
        nodes_camera = scenario.data['nodes_camera'][0]
        feats = subgraph.ndata['h']
        self.model.g = subgraph
        for layer in self.model.layers:
            layer.g = subgraph
        outputs = torch.squeeze(self.model(feats.float(), subgraph))
        indices = torch.squeeze(indices).to('cpu')
        final_output = self.smu.get_person_proposal_from_network_output(outputs, subgraph, indices, nodes_camera,
                                                                        scenario.jsons_for_head, 0.5)
        batched_input = []
        for person in final_output:                                                 # :243-274
            raw_input = {}
            for camera in p.used_cameras:
                if person[camera] is not None:
                    raw_input[camera] = [json.dumps([scenario.jsons_for_head[person[camera]]])]
            inputs = self.dataset.PoseEstimatorDataset(raw_input, p.cameras, p.joint_list, save=False)
            batched_input.append(inputs[0][0].reshape([1, inputs[0][0].size()[0]]))
        final_results = []
        if batched_input:                                                           # :278-294
            output_all = self.mlp(torch.cat(batched_input, dim=0))
            for person_id in range(output_all.shape[0]):
                results_3d = (torch.squeeze(output_all[person_id]) * 10.).to('cpu')
                x3D, y3D, z3D = results_3d[::3], results_3d[1::3], results_3d[2::3]
                final_results.append([[float(x3D[j]), float(y3D[j]), float(z3D[j])] for j in range(len(p.joint_list))])
        return final_output, final_results
