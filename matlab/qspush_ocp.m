classdef qspush_ocp < handle
    % Drop-in for the acados MATLAB class `acados_ocp` as used by NMPC_controller.m (set/get/solve/get_cost/print).
    % In NMPC_controller.create_ocp_solver (NMPC_controller.m:302-305) replace
    %     self.ocp_solver = acados_ocp(self.create_ocp_model(), self.create_ocp_opts());
    % by  self.ocp_solver = qspush_ocp(self.plant, self.Hp, self.sample_time, 'sqp');
    properties
        s; m; N;
    end
    properties (Constant)
        F = struct('constr_x0',0,'cost_y_ref',1,'cost_y_ref_e',2,'init_x',3,'init_u',4,'init_pi',5,'x',3,'u',4,'pi',5, ...
                   'cost_W',16,'constr_lh',17,'constr_uh',18,'status',32,'sqp_iter',33);
    end
    methods
        function self = qspush_ocp(plant, Hp, sample_time, nlp_solver)
            sp = plant.slider_params; flip = any(strcmp(plant.object_name, ["montana","pulirapid"]));
            self.m = qspush_mex('model_from_ply', char(sp.pcl_path), flip, 3, sp.mu_sg, sp.mu_sp, sp.m, sp.tau_max);
            self.N = Hp;
            self.s = qspush_mex('solver_create', self.m, Hp, sample_time, 1, 0, double(strcmp(nlp_solver,'sqp')));
        end
        function set(self, field, value, stage)
            if nargin < 4, stage = -1; end
            % single-stage fields: the reference passes set('cost_y_ref_e', y, Hp) (NMPC_controller.m:348); the C-ABI ignores the
            % stage of such fields as well, this keeps older libraries working
            if any(strcmp(field, {'constr_x0','cost_y_ref_e'})), stage = -1; end
            qspush_mex('set', self.s, self.F.(field), stage, double(value));
        end
        function solve(self), qspush_mex('solve', self.s); end   % the x0 wrap / rollout stay in NMPC_controller.solve
        function v = get(self, field, stage)
            switch field
                case {'status','sqp_iter'}, v = qspush_mex('get_int', self.s, self.F.(field));
                case 'time_tot', v = qspush_mex('stat', self.s, 0);
                case 'time_lin', v = qspush_mex('stat', self.s, 1);
                case 'time_qp_sol', v = qspush_mex('stat', self.s, 2);
                otherwise
                    dims = struct('u',[2 self.N],'x',[4 self.N+1],'pi',[4 self.N]); d = dims.(field);
                    if nargin < 3, v = qspush_mex('get', self.s, self.F.(field), -1, d(1), d(2));
                    else, v = qspush_mex('get', self.s, self.F.(field), stage, d(1), 1); end
            end
        end
        function c = get_cost(self), c = qspush_mex('get', self.s, 7, -1, 1, 1); end
        function print(self, varargin)
            fprintf('status %d, sqp_iter %d, residuals %s\n', self.get('status'), self.get('sqp_iter'), mat2str(qspush_mex('get', self.s, 8, -1, 4, 1)', 3));
        end
        function delete(self), qspush_mex('solver_free', self.s); qspush_mex('model_free', self.m); end
    end
end
